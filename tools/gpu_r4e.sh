#!/bin/bash
# final regression of round 2: all GPU tests, smoke, default bench; memcheck over the kernels changed late in the round
# (conv v2 / v1 / cta_group::2 / first layer, fused MMA pair + elected-lane issue in LINEAR / k_mp_pre / k_mp_gru, scaled
# 3xFP16 MP backward)
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4e_pytest.log 2>&1; tail -2 gpurun_out/r4e_pytest.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py > gpurun_out/r4e_bench.json 2> gpurun_out/r4e_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r4e_bench.err | tail -8
echo "=== L1 train step"; timeout 300 python tools/train_step_l1.py 2>/dev/null | tail -1
echo "=== memcheck"; timeout 1500 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 --log-file gpurun_out/r02_sanitize_memcheck_final.log \
  python -m pytest tests/test_gpu_conv.py tests/test_gpu_grad.py "tests/test_gpu_parity.py::test_l0_message_pass_vs_golden_and_oracle" "tests/test_gpu_parity.py::test_l1_forward_vs_golden" -m gpu -x -q -p no:cacheprovider > gpurun_out/r4e_memcheck_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r02_sanitize_memcheck_final.log; tail -2 gpurun_out/r4e_memcheck_pytest.log
