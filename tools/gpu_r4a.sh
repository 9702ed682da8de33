#!/bin/bash
mkdir -p gpurun_out
echo "=== L1 train step launch list (ncu)"; STEPS=1 timeout 500 /usr/local/cuda/bin/ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r4a_l1train_launches.csv python tools/train_step_l1.py > gpurun_out/r4a_l1train_ncu.log 2>&1; echo rc=$?; wc -l gpurun_out/r4a_l1train_launches.csv; tail -2 gpurun_out/r4a_l1train_ncu.log | cut -c1-300
