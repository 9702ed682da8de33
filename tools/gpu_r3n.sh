#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3n_pytest.log 2>&1; tail -3 gpurun_out/r3n_pytest.log
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r3n_bench.err | tail -9
echo "=== conv V=2"; SGG_CONV_V=2 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
echo "=== conv V=1"; SGG_CONV_V=1 timeout 300 python tools/conv_layers.py 2>&1 | tail -13 | cut -c1-60
echo "=== tc16_check"; timeout 300 python tools/tc16_check.py 2>&1 | grep -E "^linear M|^l1 " | cut -c1-200
