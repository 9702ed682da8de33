#!/bin/bash
mkdir -p gpurun_out
echo "=== A: grad + train-tail tests"; timeout 200 python -m pytest tests/test_gpu_grad.py tests/test_gpu_train_tail.py -m gpu -x -q -o faulthandler_timeout=100 > gpurun_out/a1_grad.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/a1_grad.log | cut -c1-250
echo "=== D: train step"; timeout 100 python tools/train_step.py > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; tail -3 gpurun_out/train_step.err; cat gpurun_out/train_step.json
