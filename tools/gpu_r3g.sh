#!/bin/bash
echo "=== L01 dbg"; SGG_CONV_DBG=1 SGG_CONV_V=2 SGG_CONV_CG=1 CL=1 CB=8 CREPS=1 timeout 100 python tools/conv_one.py 2>&1 | grep "conv dbg" | head -14
echo "=== L08 dbg"; SGG_CONV_DBG=1 SGG_CONV_V=2 SGG_CONV_CG=1 CL=8 CB=8 CREPS=1 timeout 100 python tools/conv_one.py 2>&1 | grep "conv dbg" | head -14
