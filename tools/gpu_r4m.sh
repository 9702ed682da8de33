#!/bin/bash
# end-of-round regression: all GPU tests, smoke, default bench, reference arm, memcheck over the kernels added last
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4m_pytest.log 2>&1; tail -2 gpurun_out/r4m_pytest.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py > gpurun_out/r4m_bench.json 2> gpurun_out/r4m_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r4m_bench.err | tail -8
echo "=== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r4m_ref.json 2> gpurun_out/r4m_ref.err; cut -c1-200 gpurun_out/r4m_ref.json
echo "=== memcheck"; timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 --log-file gpurun_out/r02_sanitize_memcheck_lin16p.log \
  python -m pytest tests/test_gpu_lin16p.py tests/test_gpu_train_ops.py "tests/test_gpu_model.py" -m gpu -x -q -p no:cacheprovider > gpurun_out/r4m_memcheck_pytest.log 2>&1; echo rc=$?; tail -2 gpurun_out/r02_sanitize_memcheck_lin16p.log; tail -2 gpurun_out/r4m_memcheck_pytest.log
