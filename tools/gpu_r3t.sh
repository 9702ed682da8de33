#!/bin/bash
for L in 1 3; do
echo "=== conv_one L$L CSECS=3"; SGG_CONV_V=2 CL=$L CB=32 CSECS=3 timeout 100 python tools/conv_one.py 2>&1 | grep -E "launches|clock samples" | cut -c1-700
done
