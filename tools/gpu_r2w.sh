#!/bin/bash
mkdir -p gpurun_out
for cg in 2 1; do
echo "=== SGG_CONV_CG=$cg"
SGG_CONV_CG=$cg timeout 300 python tools/conv_check.py > gpurun_out/r2w_conv_cg$cg.log 2>&1; echo rc=$?; grep -v "^layers" gpurun_out/r2w_conv_cg$cg.log | tail -30
done
echo "=== pytest conv"; timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q 2>&1 | tail -3
