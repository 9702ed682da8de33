#!/bin/bash
# Round checkpoint: full GPU test suite, bench (both arms), smoke, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
echo "=== A: pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=150 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_pytest.log
echo "=== B: reference arm"; timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c_ref.json 2> gpurun_out/c_ref.err; echo "rc=$?"; cut -c1-200 gpurun_out/c_ref.json
echo "=== C: bench"; timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "rc=$?"; tail -4 gpurun_out/b_bench.err; cut -c1-200 gpurun_out/b_bench.json
echo "=== D: smoke"; timeout 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/f_smoke.log
echo "=== E: ncu launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/d_ncu.log 2>&1; echo "rc=$?"
echo "=== F: ncu full (edge gru + edge unary)"; timeout 250 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_tc16<\(int\)3, \(int\)80, \(int\)1, \(int\)3|k_tc16<\(int\)1, \(int\)128" -s 1 -c 3 -o gpurun_out/e_prof python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/e_ncu.log 2>&1; echo "rc=$?"
echo "=== G2: cfg3-shaped runs (B=32): L1 bench line + feature-head stage timings"; timeout 200 python bench.py --batch 32 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench_b32.json 2>/dev/null; cut -c1-160 gpurun_out/b_bench_b32.json; timeout 200 python tools/bench_l2.py > gpurun_out/bench_l2.json 2>&1; cat gpurun_out/bench_l2.json | tail -1 | cut -c1-900
echo "=== G: phases"; SGG_TC_TIMING=1 timeout 120 python tools/tc16_phases.py > gpurun_out/phases.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -30
