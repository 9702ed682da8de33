#!/bin/bash
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
echo "=== clocks under a 4 s conv loop (layer 8, CG=1)"; SGG_CONV_CG=1 CL=8 CSECS=4 timeout 120 python tools/conv_one.py 2>&1 | tail -3
for cg in 1 2; do
echo "=== ncu layer 8 CG=$cg"
SGG_CONV_CG=$cg CL=8 timeout 250 $NCU --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 1 -c 1 -f -o gpurun_out/r2y_conv_cg$cg python tools/conv_one.py > gpurun_out/r2y_ncu_cg$cg.log 2>&1; echo rc=$?
done
echo "=== ncu layer 1 CG=1"
SGG_CONV_CG=1 CL=1 CB=8 timeout 250 $NCU --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 1 -c 1 -f -o gpurun_out/r2y_conv_l1 python tools/conv_one.py > gpurun_out/r2y_ncu_l1.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
