#!/bin/bash
mkdir -p gpurun_out
echo "=== A: pytest gpu"; timeout 400 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/a_pytest.log
echo "=== B: bench"; timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "rc=$?"; cat gpurun_out/b_bench.err | tail -8; cut -c1-300 gpurun_out/b_bench.json
echo "=== C: probe"; SGG_CHECK_SKIP_LINEAR=1 SGG_CHECK_MODES=simt,tc16 timeout 150 python tools/tc16_check.py > gpurun_out/c_check.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/c_check.log
echo "=== D: ncu launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/d_ncu.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/d_launches.csv run | head -30
