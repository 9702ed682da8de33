#!/bin/bash
for L in 1 2; do
echo "=== conv_one L$L CEACH"; SGG_CONV_V=2 CL=$L CB=32 CEACH=1 timeout 100 python tools/conv_one.py 2>&1 | tail -4
done
echo "=== ncu L1 v2 B=32"; SGG_CONV_V=2 CL=1 CB=32 timeout 200 /usr/local/cuda/bin/ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_conv3x3 python tools/conv_one.py 2>&1 | grep -E "gpu__time|per_second|dram__" | head -12
