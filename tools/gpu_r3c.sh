#!/bin/bash
for kcb in 4 8 16 100000; do
echo "=== V=2 CG=1 KCB=$kcb"
SGG_CONV_KCB=$kcb SGG_CONV_V=2 SGG_CONV_CG=1 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
done
echo "=== V=2 CG=2 KCB=100000"
SGG_CONV_KCB=100000 SGG_CONV_V=2 SGG_CONV_CG=2 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
