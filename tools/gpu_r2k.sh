#!/bin/bash
# 2-GPU validation of the bench (L1 forward replicas + train step with NCCL gradient all-reduce)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2k_topo.txt 2>&1
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k union_geom 2>&1 | tail -2
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL SGG_BENCH_WATCHDOG=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo rc=$?
grep -E "^\[bench" gpurun_out/r2k_bench.err | tail -14
grep -E "NVLS|Using network|comm .* rank|Channel.*via|Connected all" gpurun_out/r2k_bench.err | head -8
python - <<'PY'
import json
try:
    lines = [l for l in open('gpurun_out/r2k_bench.json').read().strip().splitlines() if l.startswith('{')]
    d = json.loads(lines[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e'])
    print(json.dumps(d.get('train_step'))[400:2400])
except Exception as e:
    print('no json', e)
PY
