#!/bin/bash
mkdir -p gpurun_out
echo "=== A: pytest gpu"; timeout 400 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=120 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_pytest.log
echo "=== B: bench"; timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 >/dev/null | grep -E "device-resident|stage"
echo "=== B2: bench, stream-K off"; SGG_TC16_STREAMK=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 >/dev/null | grep -E "device-resident|stage"
echo "=== C: probe"; SGG_CHECK_MODES=tc16 timeout 150 python tools/tc16_check.py 2>&1 | grep linear
echo "=== D: ncu launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/d_ncu.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/d_launches.csv run | head -16
