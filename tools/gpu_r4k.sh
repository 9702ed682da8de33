#!/bin/bash
echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_lin16p.py tests/test_gpu_model.py tests/test_gpu_train_ops.py tests/test_gpu_grad.py -m gpu -x -q 2>&1 | tail -2
echo "=== train step L1"; timeout 300 python tools/train_step_l1.py 2>/dev/null | tail -1
echo "=== train step"; timeout 300 python tools/train_step.py 2>/dev/null | tail -1
