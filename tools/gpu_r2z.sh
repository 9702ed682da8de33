#!/bin/bash
mkdir -p gpurun_out
echo "=== umma shift"; timeout 120 tools/ubench/umma_shift 2>&1 | tail -10
echo "=== train step launch list (ncu)"; STEPS=1 timeout 500 /usr/local/cuda/bin/ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2z_train_launches.csv python tools/train_step.py > gpurun_out/r2z_train_ncu.log 2>&1; echo rc=$?; wc -l gpurun_out/r2z_train_launches.csv
