#!/bin/bash
# 4-GPU check of the bench (what the driver's SCALE run does at N = 4) + smoke on GPU 0
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SGG_BENCH_WATCHDOG=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo rc=$?
grep -E "^\[bench" gpurun_out/r2r_bench.err | tail -10
python - <<'PY'
import json
try:
    lines = [l for l in open('gpurun_out/r2r_bench.json').read().strip().splitlines() if l.startswith('{')]
    d = json.loads(lines[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, {k: d['e2e'][k] for k in ('value', 'h2d_gbs_per_rank', 'h2d_plain_memcpy_gbs_per_rank')})
    t = d['train_step']; print({k: t.get(k) for k in ('ms_per_step', 'images_per_s', 'ms_per_step_no_collectives', 'exposed_comm_ms', 'allreduce_standalone', 'comm_hidden_frac')}, t.get('e2e', {}).get('images_per_s'))
except Exception as e:
    print('no json', e)
PY
