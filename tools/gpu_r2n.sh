#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_check.py > gpurun_out/r2n_conv.log 2>&1; echo rc=$?; tail -40 gpurun_out/r2n_conv.log
