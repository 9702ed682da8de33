#!/bin/bash
# backward / RoIAlign iteration: GPU tests, feature-head stages, train step
mkdir -p gpurun_out
echo "=== A0: new tests"; timeout 300 python -m pytest tests/test_gpu_grad.py tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core_backward or roi_align or gradients" -o faulthandler_timeout=120 > gpurun_out/a0_new.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/a0_new.log | cut -c1-250
echo "=== A: pytest gpu (all)"; timeout 500 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=150 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/a_pytest.log | cut -c1-220
echo "=== C: feature-head stages (B=32)"; timeout 200 python tools/bench_l2.py > gpurun_out/bench_l2.json 2>&1; tail -1 gpurun_out/bench_l2.json | cut -c1-800
echo "=== D: train step"; timeout 200 python tools/train_step.py > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; tail -3 gpurun_out/train_step.err; cat gpurun_out/train_step.json
echo "=== D2: train step, SIMT backward"; SGG_BWD_TC=0 timeout 200 python tools/train_step.py 2>/dev/null | cut -c100-260
