#!/bin/bash
echo "=== grad tests"; timeout 600 python -m pytest tests/test_gpu_grad.py -m gpu -x -q 2>&1 | tail -2
echo "=== train step L1 + profile"; PROFILE=1 timeout 300 python tools/train_step_l1.py 2>&1 | tail -18 | cut -c1-220
