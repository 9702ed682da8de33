#!/bin/bash
# backward / RoIAlign iteration: GPU tests, feature-head stages, train step
mkdir -p gpurun_out
echo "=== A: pytest gpu (all)"; timeout 500 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=150 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/a_pytest.log | cut -c1-220
echo "=== C: feature-head stages (B=32)"; timeout 200 python tools/bench_l2.py > gpurun_out/bench_l2.json 2>&1; tail -1 gpurun_out/bench_l2.json | cut -c1-700
echo "=== D: train step"; timeout 200 python tools/train_step.py > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; tail -3 gpurun_out/train_step.err; cat gpurun_out/train_step.json
