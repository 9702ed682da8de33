#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/relu_flip_check.py 2>&1 | tail -5
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; tail -8 gpurun_out/r2j_pytest.log
echo "=== train step"; timeout 300 python tools/train_step.py > gpurun_out/r2j_train.json 2> gpurun_out/r2j_train.err; echo rc=$?; tail -3 gpurun_out/r2j_train.err; cat gpurun_out/r2j_train.json
