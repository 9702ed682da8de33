#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest train ops + grad"; timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_grad.py tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -5 gpurun_out/r2q_pytest.log
echo "=== train step tc16 bwd"; timeout 300 python tools/train_step.py 2> /dev/null | tail -1
echo "=== train step tc32 bwd"; SGG_BWD_ENGINE=tc32 timeout 300 python tools/train_step.py 2> /dev/null | tail -1
echo "=== L1 train step"; timeout 300 python tools/train_step_l1.py 2>/dev/null | tail -1
echo "=== memcheck (new kernels)"; timeout 1200 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 --log-file gpurun_out/sanitize_memcheck_r2b.log \
  python -m pytest "tests/test_gpu_conv.py::test_conv_stack_prefixes_vs_cudnn_fp32" tests/test_gpu_train_ops.py "tests/test_gpu_parity.py::test_union_geom_train_mode_forward_backward_vs_reference" "tests/test_gpu_parity.py::test_l0_message_pass_vs_golden_and_oracle" "tests/test_gpu_parity.py::test_l1_forward_vs_golden" -m gpu -x -q -p no:cacheprovider > gpurun_out/sanitize_memcheck_r2b.pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/sanitize_memcheck_r2b.log; tail -2 gpurun_out/sanitize_memcheck_r2b.pytest.log
