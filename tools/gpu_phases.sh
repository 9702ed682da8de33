#!/bin/bash
mkdir -p gpurun_out
SGG_TC_TIMING=1 timeout 120 python tools/tc16_phases.py > gpurun_out/phases.log 2>&1; echo "rc=$?"; cat gpurun_out/phases.log
