#!/bin/bash
# quick validation: GPU tests, bench (no CPU baseline), engine probe
mkdir -p gpurun_out
echo "=== A: pytest gpu"; timeout 400 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/a_pytest.log
echo "=== B: bench"; timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 >gpurun_out/b_bench.json | grep -E "device-resident|e2e|stage"
echo "=== B32: bench"; timeout 300 python bench.py --batch 32 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 >/dev/null | grep -E "device-resident"
echo "=== C: probe"; SGG_CHECK_MODES=tc16 timeout 150 python tools/tc16_check.py 2>&1 | grep -E "linear M= (2400|9600| 1000|  240 N=  512)"
