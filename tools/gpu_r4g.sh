#!/bin/bash
mkdir -p gpurun_out
echo "=== new tests"; timeout 600 python -m pytest tests/test_gpu_lin16p.py tests/test_gpu_model.py tests/test_gpu_eval_tail.py -m gpu -x -q 2>&1 | tail -8
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --no-train > gpurun_out/r4g_bench.json 2> gpurun_out/r4g_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r4g_bench.err | tail -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4g_bench.json').read().strip().splitlines()[-1])
oc=d['config'].get('other_configs') or d.get('other_configs')
print(json.dumps(oc.get('cfg3_feature_head'),indent=0)[:900]); print({k:oc['cfg3_l3_forward_e2e'].get(k) for k in ('ms_per_batch','images_per_s','backbone_ms')})
PY
