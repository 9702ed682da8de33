#!/bin/bash
mkdir -p gpurun_out
for cg in 1 2; do
echo "=== SGG_CONV_CG=$cg"
SGG_CONV_CG=$cg timeout 300 python tools/conv_layers.py 2>&1 | tail -14
done
SGG_CONV_CG=2 timeout 300 python tools/conv_check.py 2>&1 | tail -3
