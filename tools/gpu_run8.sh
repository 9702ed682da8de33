#!/bin/bash
mkdir -p gpurun_out
echo "=== A: pytest gpu"; timeout 400 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_pytest.log
echo "=== B: bench"; timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "rc=$?"; cat gpurun_out/b_bench.err | tail -5
echo "=== C: probe"; SGG_CHECK_SKIP_LINEAR=1 SGG_CHECK_MODES=simt,tc16 timeout 150 python tools/tc16_check.py > gpurun_out/c_check.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/c_check.log
echo "=== D: phases"; SGG_TC_TIMING=1 timeout 120 python tools/tc16_phases.py > gpurun_out/phases.log 2>&1; echo "rc=$?"; cat gpurun_out/phases.log
