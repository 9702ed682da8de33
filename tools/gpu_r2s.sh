#!/bin/bash
mkdir -p gpurun_out
for d in 1 0; do
echo "=== DIRECT=$d tc16_check"; SGG_TC16_DIRECT=$d SGG_CHECK_MODES=tc16 timeout 200 python tools/tc16_check.py 2>&1 | grep -E "linear M|l1 B" | tail -16
echo "=== DIRECT=$d l1"; SGG_TC16_DIRECT=$d timeout 200 python tools/mpf_check.py time 2>&1 | grep -E "l1 cuda-graph|message_pass T=3"
echo "=== DIRECT=$d train"; SGG_TC16_DIRECT=$d timeout 300 python tools/train_step.py 2>/dev/null | tail -1
done
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; tail -4 gpurun_out/r2s_pytest.log
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; grep -E "^\[bench" gpurun_out/r2s_bench.err | tail -8
