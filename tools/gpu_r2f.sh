#!/bin/bash
mkdir -p gpurun_out
timeout 250 python tools/mpf_check.py all > gpurun_out/r2f_a.log 2>&1; tail -9 gpurun_out/r2f_a.log
SGG_CHECK_MODES=tc16 timeout 200 python tools/tc16_check.py > gpurun_out/r2f_tc16check.log 2>&1; tail -30 gpurun_out/r2f_tc16check.log
timeout 300 python tools/bench_l2.py > gpurun_out/r2f_l2.json 2> gpurun_out/r2f_l2.err; cut -c1-700 gpurun_out/r2f_l2.json
echo "=== bench"; timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err; cut -c1-300 gpurun_out/r2f_bench.json
