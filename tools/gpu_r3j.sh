#!/bin/bash
for cfg in "1 100000" "1 4" "0 100000" "0 4" "2 4"; do
set -- $cfg
echo "=== V=2 CG=1 DRY=$1 KCB=$2"
SGG_CONV_DRY=$1 SGG_CONV_KCB=$2 SGG_CONV_V=2 SGG_CONV_CG=1 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
done
