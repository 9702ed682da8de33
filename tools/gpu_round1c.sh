#!/bin/bash
# eval-tail validation + evidence for the training tail: tests, smoke, ncu capture of the SGD sweep, train-step launch list
mkdir -p gpurun_out
echo "=== A0: eval-tail + model tests"; timeout 200 python -m pytest tests/test_gpu_eval_tail.py tests/test_gpu_model.py -m gpu -q -o faulthandler_timeout=100 > gpurun_out/a0_eval_tail.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/a0_eval_tail.log | cut -c1-250
echo "=== S: smoke"; timeout 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/f_smoke.log | cut -c1-250
echo "=== A: pytest gpu (all)"; timeout 400 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=150 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_pytest.log | cut -c1-220
echo "=== N: ncu full, SGD sweep"; OPTIM_BENCH_SKIP_REF=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_mt_sgd -s 4 -c 1 -f -o gpurun_out/r01_mt_sgd python tools/optim_bench.py > gpurun_out/n_ncu_sgd.log 2>&1; echo "rc=$?"
echo "=== T: train-step launch list"; STEPS=1 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_step.py > gpurun_out/train_ncu.log 2>&1; echo "rc=$?"
ls -la gpurun_out | grep -E "r01_mt_sgd|train_launches"
