#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3w_pytest.log 2>&1; tail -2 gpurun_out/r3w_pytest.log
echo "=== stack"; timeout 200 python tools/conv_stack_events.py 2>&1 | tail -1
timeout 300 python tools/conv_check.py 2>&1 | grep -v "^layers" | tail -2
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r3w_bench.json 2> gpurun_out/r3w_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r3w_bench.err | tail -9
