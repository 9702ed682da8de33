#!/bin/bash
for dry in 2 3 4; do
echo "=== V=2 CG=1 DRY=$dry"
SGG_CONV_DRY=$dry SGG_CONV_V=2 SGG_CONV_CG=1 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60 | head -4
done
