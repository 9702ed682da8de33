#!/bin/bash
mkdir -p gpurun_out
echo "=== A: pytest gpu"; timeout 400 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_pytest.log
echo "=== B: bench"; timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "rc=$?"; cat gpurun_out/b_bench.err | tail -4
echo "=== B2: bench, stream-K off"; SGG_TC16_STREAMK=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 >/dev/null | grep -E "device-resident|stage"
echo "=== C: probe"; SGG_CHECK_MODES=simt,tc16 timeout 150 python tools/tc16_check.py > gpurun_out/c_check.log 2>&1; echo "rc=$?"; tail -22 gpurun_out/c_check.log
echo "=== D: phases"; SGG_TC_TIMING=1 timeout 120 python tools/tc16_phases.py > gpurun_out/phases.log 2>&1; echo "rc=$?"; head -3 gpurun_out/phases.log
