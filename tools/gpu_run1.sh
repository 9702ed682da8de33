#!/bin/bash
# GPU session 1: baseline (3xTF32 default) tests + bench, then the 3xFP16 engine probe / tests / bench / profiles.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== A: pytest gpu (default engine)"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest_tc32.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_pytest_tc32.log
echo "=== B: bench (default engine)"; timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/b_bench_tc32.json 2> gpurun_out/b_bench_tc32.err; echo "rc=$?"; cut -c1-600 gpurun_out/b_bench_tc32.json
echo "=== C: tc16 probe"; SGG_TC_MODE=1 timeout 150 python tools/tc16_check.py > gpurun_out/c_tc16_check.log 2>&1; rc=$?; echo "rc=$rc"; tail -40 gpurun_out/c_tc16_check.log
if [ $rc -ne 0 ]; then echo "tc16 probe failed: skipping D-G"; ls -la gpurun_out; exit 0; fi
echo "=== D: pytest gpu (tc16)"; SGG_TC_MODE=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/d_pytest_tc16.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/d_pytest_tc16.log
echo "=== E: bench (tc16)"; SGG_TC_MODE=1 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_tc16.json 2> gpurun_out/e_bench_tc16.err; echo "rc=$?"; cut -c1-600 gpurun_out/e_bench_tc16.json
echo "=== F: ncu launch list (tc16)"; SGG_TC_MODE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/f_launches_tc16.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/f_ncu.log 2>&1; echo "rc=$?"
echo "=== G: ncu full (tc16 edge gru)"; SGG_TC_MODE=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:\\(int\\)80, \\(int\\)1, \\(int\\)3" -s 1 -c 2 -o gpurun_out/g_prof_tc16 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/g_ncu.log 2>&1; echo "rc=$?"
ls -la gpurun_out
