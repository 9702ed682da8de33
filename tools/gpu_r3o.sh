#!/bin/bash
for L in 1 2 8; do
echo "=== L$L dbg B=32"; SGG_CONV_DBG=1 SGG_CONV_V=2 CL=$L CB=32 CREPS=2 timeout 100 python tools/conv_one.py 2>&1 | grep "conv dbg" | tail -4
done
echo "=== L1 dbg B=32 DRY=1"; SGG_CONV_DRY=1 SGG_CONV_DBG=1 SGG_CONV_V=2 CL=1 CB=32 CREPS=2 timeout 100 python tools/conv_one.py 2>&1 | grep "conv dbg" | tail -3
