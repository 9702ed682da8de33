#!/bin/bash
mkdir -p gpurun_out
echo "=== racecheck (SIMT shared-memory kernels: loss / clip / SGD / ranking / BatchNorm / pooling / RoIAlign)"
SGG_GEMM=simt timeout 800 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --log-file gpurun_out/r02_sanitize_racecheck.log \
  python -m pytest tests/test_gpu_train_tail.py tests/test_gpu_eval_tail.py tests/test_gpu_train_ops.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r4q_racecheck_pytest.log 2>&1; echo rc=$?
tail -3 gpurun_out/r02_sanitize_racecheck.log; tail -2 gpurun_out/r4q_racecheck_pytest.log
grep -c "Race reported\|hazard" gpurun_out/r02_sanitize_racecheck.log
