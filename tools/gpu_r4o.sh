#!/bin/bash
mkdir -p gpurun_out
STEPS=1 timeout 500 /usr/local/cuda/bin/ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_train_final.csv python tools/train_step.py > gpurun_out/r4o_train_ncu.log 2>&1; echo rc=$?; wc -l gpurun_out/r02_launches_train_final.csv
