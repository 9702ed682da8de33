#!/bin/bash
mkdir -p gpurun_out
for c in 3.0 0.5; do
echo "=== split cost $c"; SGG_TC16_SPLIT_COST=$c timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 >/dev/null | grep -E "device-resident|e2e|stage"
done
echo "=== ncu launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/d_ncu.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/d_launches.csv run | head -28
