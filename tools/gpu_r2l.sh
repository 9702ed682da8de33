#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; tail -8 gpurun_out/r2l_pytest.log
echo "=== train step"; timeout 300 python tools/train_step.py > gpurun_out/r2l_train.json 2> gpurun_out/r2l_train.err; echo rc=$?; tail -3 gpurun_out/r2l_train.err; cat gpurun_out/r2l_train.json
echo "=== train step L1"; timeout 300 python tools/train_step_l1.py > gpurun_out/r2l_train_l1.json 2> gpurun_out/r2l_train_l1.err; echo rc=$?; tail -3 gpurun_out/r2l_train_l1.err; cat gpurun_out/r2l_train_l1.json
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r2l_bench.err | tail -12
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches_per_step')}, d['e2e']['value'], d['cpu_baseline'])
    r = d['roofline']; print({k: r[k] for k in ('bound', 'achieved', 'peak', 'frac', 'ms_per_launch', 'hbm', 'stage_ms')})
    print(json.dumps(d.get('train_step'))[400:1500])
except Exception as e:
    print('no json', e)
PY
