#!/bin/bash
for L in 1 8; do
echo "=== L$L dbg B=32"; SGG_CONV_DBG=1 SGG_CONV_V=2 CL=$L CB=32 CREPS=2 timeout 100 python tools/conv_one.py 2>&1 | grep "conv dbg" | tail -2
done
