#!/bin/bash
mkdir -p gpurun_out
echo "=== bench (train nested)"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo rc=$?; tail -25 gpurun_out/r2g_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches_per_step')}); print(json.dumps(d.get('train_step'))[:1500])
except Exception as e:
    print('no json', e)
PY
