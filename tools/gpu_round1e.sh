#!/bin/bash
mkdir -p gpurun_out
echo "=== A: pytest gpu (all)"; timeout 150 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=100 > gpurun_out/a_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/a_pytest.log | cut -c1-200
echo "=== B5: cfg5-shaped bench (8 img, 64 boxes, 2000 edges, T=6)"; timeout 80 python bench.py --batch 8 --boxes 64 --edges 2000 --iters 6 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench_cfg5.json 2> gpurun_out/b_bench_cfg5.err; echo "rc=$?"; cut -c1-330 gpurun_out/b_bench_cfg5.json
