#!/bin/bash
echo "=== conv V=2 CG=1 (elect-issue)"; SGG_CONV_V=2 SGG_CONV_CG=1 timeout 300 python tools/conv_check.py 2>&1 | grep -v "^layers" | tail -4
SGG_CONV_V=2 SGG_CONV_CG=1 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
echo "=== conv V=1 CG=1"; SGG_CONV_V=1 SGG_CONV_CG=1 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
echo "=== L01 dbg"; SGG_CONV_DBG=1 SGG_CONV_V=2 SGG_CONV_CG=1 CL=1 CB=8 CREPS=1 timeout 100 python tools/conv_one.py 2>&1 | grep "conv dbg" | head -6
echo "=== L08 dbg"; SGG_CONV_DBG=1 SGG_CONV_V=2 SGG_CONV_CG=1 CL=8 CB=8 CREPS=1 timeout 100 python tools/conv_one.py 2>&1 | grep "conv dbg" | head -6
