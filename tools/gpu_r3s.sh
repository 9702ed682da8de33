#!/bin/bash
CEACH=1 SGG_CONV_V=2 timeout 200 python tools/conv_layers.py 2>&1 | head -8
