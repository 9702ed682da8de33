#!/bin/bash
# GPU session: full test suite (default engine 3xFP16; parity tests cover tc16/tc32/simt), bench, reference arm, profiles.
mkdir -p gpurun_out
echo "=== A: pytest gpu"; timeout 500 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=150 --durations=8 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/a_pytest.log
echo "=== B: bench"; timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "rc=$?"; cat gpurun_out/b_bench.err | tail -12; cut -c1-1500 gpurun_out/b_bench.json
echo "=== C: reference arm"; timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c_ref.json 2> gpurun_out/c_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/c_ref.json
echo "=== D: ncu launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/d_ncu.log 2>&1; echo "rc=$?"
echo "=== E: ncu full (edge gru)"; timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:\(int\)80, \(int\)1, \(int\)3" -s 1 -c 2 -o gpurun_out/e_prof_edge python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/e_ncu.log 2>&1; echo "rc=$?"
echo "=== F: smoke"; timeout 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/f_smoke.log
echo "=== G: tc32 hang probe"; timeout 300 python tools/hang_probe.py tc32 > gpurun_out/g_hang_tc32.log 2>&1; echo "rc=$?"; cat gpurun_out/g_hang_tc32.log
ls -la gpurun_out
