#!/bin/bash
for dry in 1 2 0; do
echo "=== V=2 CG=1 DRY=$dry"
SGG_CONV_DRY=$dry SGG_CONV_V=2 SGG_CONV_CG=1 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
done
echo "=== V=2 CG=2 DRY=1"
SGG_CONV_DRY=1 SGG_CONV_V=2 SGG_CONV_CG=2 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
echo "=== V=2 CG=2 DRY=2"
SGG_CONV_DRY=2 SGG_CONV_V=2 SGG_CONV_CG=2 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
