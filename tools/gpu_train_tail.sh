#!/bin/bash
# training-tail validation: new GPU tests first, then the whole GPU suite, optimizer / feature-head / train-step timings, bench sanity
mkdir -p gpurun_out
echo "=== A0: train-tail tests"; timeout 300 python -m pytest tests/test_gpu_train_tail.py -m gpu -q -o faulthandler_timeout=120 > gpurun_out/a0_train_tail.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/a0_train_tail.log | cut -c1-220
echo "=== A: pytest gpu (all)"; timeout 500 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=150 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_pytest.log | cut -c1-220
echo "=== B: optimizer sweeps"; timeout 200 python tools/optim_bench.py > gpurun_out/optim_bench.json 2> gpurun_out/optim_bench.err; echo "rc=$?"; tail -3 gpurun_out/optim_bench.err; cat gpurun_out/optim_bench.json
echo "=== C: feature-head stages (B=32)"; timeout 200 python tools/bench_l2.py > gpurun_out/bench_l2.json 2>&1; tail -1 gpurun_out/bench_l2.json | cut -c1-1200
echo "=== D: train step"; timeout 200 python tools/train_step.py > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; tail -3 gpurun_out/train_step.err; cat gpurun_out/train_step.json
echo "=== E: bench"; timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/b_bench.err >gpurun_out/b_bench.json; echo "rc=$?"; cut -c1-260 gpurun_out/b_bench.json
