#!/bin/bash
# per-kernel durations of one PredCls L1 train step (B=32): where do the 12 ms go?
mkdir -p gpurun_out
STEPS=1 timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_step.py > gpurun_out/train_ncu.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/train_launches.csv 2>&1 | head -60
