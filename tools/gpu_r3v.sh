#!/bin/bash
echo "=== pytest conv + smoke"; timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -2
echo "=== stack"; timeout 200 python tools/conv_stack_events.py 2>&1 | tail -1
timeout 300 python tools/conv_check.py 2>&1 | grep -v "^layers" | tail -3
