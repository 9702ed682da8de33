#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; tail -4 gpurun_out/r2p_pytest.log
echo "=== memcheck (new kernels)"; timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 --log-file gpurun_out/sanitize_memcheck_r2b.log \
  python -m pytest "tests/test_gpu_conv.py::test_conv_stack_prefixes_vs_cudnn_fp32" tests/test_gpu_train_ops.py "tests/test_gpu_parity.py::test_union_geom_train_mode_forward_backward_vs_reference" "tests/test_gpu_parity.py::test_l0_message_pass_vs_golden_and_oracle" "tests/test_gpu_parity.py::test_l1_forward_vs_golden_and_oracle" -m gpu -x -q -p no:cacheprovider > gpurun_out/sanitize_memcheck_r2b.pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/sanitize_memcheck_r2b.log; tail -2 gpurun_out/sanitize_memcheck_r2b.pytest.log
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r2p_bench.err | tail -8
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches_per_step')}, d['e2e']['value'])
print(json.dumps(d['other_configs'])[:2500])
PY
