#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; tail -4 gpurun_out/r2u_pytest.log
echo "=== bench"; SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo rc=$?; grep -E "^\[bench" gpurun_out/r2u_bench.err | tail -9
echo "=== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2u_ref.json 2> gpurun_out/r2u_ref.err; cut -c1-400 gpurun_out/r2u_ref.json
