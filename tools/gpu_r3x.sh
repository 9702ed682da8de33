#!/bin/bash
# Round-2 final evidence: tests, launch list of one L1 step, ncu --set full captures of the fused MP kernels, the edge-unary
# LINEAR kernel and two conv layers, default bench run.
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3x_pytest.log 2>&1; tail -2 gpurun_out/r3x_pytest.log
echo "=== launch list"; timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-train --no-other-configs --no-cpu-baseline > gpurun_out/r3x_launches_bench.log 2>&1; echo rc=$?
echo "=== ncu gru"; timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_mp_gru -s 6 -c 2 -o gpurun_out/r02f_prof_gru -f \
  python tools/mpf_check.py time > gpurun_out/r3x_prof_gru.log 2>&1; echo rc=$?
echo "=== ncu pre"; timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_mp_pre -s 4 -c 1 -o gpurun_out/r02f_prof_pre -f \
  python tools/mpf_check.py time > gpurun_out/r3x_prof_pre.log 2>&1; echo rc=$?
echo "=== ncu lin"; timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_tc16 -s 2 -c 1 -o gpurun_out/r02f_prof_lin -f \
  python tools/mpf_check.py time > gpurun_out/r3x_prof_lin.log 2>&1; echo rc=$?
echo "=== ncu conv L8 v2"; CL=8 timeout 250 $NCU --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 1 -c 1 -f -o gpurun_out/r02f_conv_l8 python tools/conv_one.py > gpurun_out/r3x_ncu_l8.log 2>&1; echo rc=$?
echo "=== ncu conv L1 v2"; CL=1 timeout 250 $NCU --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 1 -c 1 -f -o gpurun_out/r02f_conv_l1 python tools/conv_one.py > gpurun_out/r3x_ncu_l1.log 2>&1; echo rc=$?
ls -la gpurun_out/r02f*.ncu-rep
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py > gpurun_out/r3x_bench.json 2> gpurun_out/r3x_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r3x_bench.err | tail -9
echo "=== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r3x_ref.json 2> gpurun_out/r3x_ref.err; cut -c1-300 gpurun_out/r3x_ref.json
