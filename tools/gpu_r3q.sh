#!/bin/bash
echo "=== conv_one L1 CSECS=1"; SGG_CONV_V=2 CL=1 CB=32 CSECS=1 timeout 100 python tools/conv_one.py 2>&1 | tail -2 | cut -c1-300
echo "=== conv_layers with DBG (first layers)"; SGG_CONV_DBG=1 SGG_CONV_V=2 timeout 200 python tools/conv_layers.py 2>&1 | grep -E "CTAs: span|^L0[12]" | head -10
