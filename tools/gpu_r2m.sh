#!/bin/bash
# ncu evidence for round 2: launch list of one L1 step (eager launches) + full captures of the fused MP kernels
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-train --no-other-configs --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1; echo rc=$?
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_mp_gru -s 6 -c 2 -o gpurun_out/r02_prof_gru -f \
  python tools/mpf_check.py time > gpurun_out/r02_prof_gru.log 2>&1; echo rc=$?
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_mp_pre -s 4 -c 1 -o gpurun_out/r02_prof_pre -f \
  python tools/mpf_check.py time > gpurun_out/r02_prof_pre.log 2>&1; echo rc=$?
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:k_tc16 -s 2 -c 1 -o gpurun_out/r02_prof_lin -f \
  python tools/mpf_check.py time > gpurun_out/r02_prof_lin.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
echo "=== bench (w96 + h2d ceiling)"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --steps 50 --warmup 5 --no-train --no-other-configs --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; grep -E "^\[bench" gpurun_out/r2m_bench.err | tail -6; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches_per_step')}, d['e2e'])
r = d['roofline']; print({k: r[k] for k in ('bound', 'achieved', 'peak', 'frac', 'ms_per_launch', 'stage_ms')})
PY
echo "=== w96 off"; SGG_TC16_W96=0 timeout 200 python tools/mpf_check.py time 2>&1 | tail -4
echo "=== w96 on"; timeout 200 python tools/mpf_check.py time 2>&1 | tail -4
echo "=== pytest parity"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
