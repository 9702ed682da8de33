#!/bin/bash
mkdir -p gpurun_out
echo "=== umma shift"; timeout 120 tools/ubench/umma_shift 2>&1 | tail -12
for cfg in "2 1" "2 2"; do
set -- $cfg
echo "=== SGG_CONV_V=$1 SGG_CONV_CG=$2"
SGG_CONV_V=$1 SGG_CONV_CG=$2 timeout 300 python tools/conv_check.py > gpurun_out/r3b_conv_v$1_cg$2.log 2>&1; echo rc=$?; grep -v "^layers" gpurun_out/r3b_conv_v$1_cg$2.log | tail -29
SGG_CONV_V=$1 SGG_CONV_CG=$2 timeout 300 python tools/conv_layers.py 2>&1 | tail -14
done
