#!/bin/bash
mkdir -p gpurun_out
env SGG_TC_TIMING=1 timeout 250 python tools/mpf_check.py all > gpurun_out/r2e_a.log 2>&1; tail -22 gpurun_out/r2e_a.log
env SGG_TC16_PF=0 timeout 200 python tools/mpf_check.py time > gpurun_out/r2e_pf0.log 2>&1; tail -8 gpurun_out/r2e_pf0.log
env SGG_TC16_PF=16 timeout 200 python tools/mpf_check.py time > gpurun_out/r2e_pf16.log 2>&1; tail -8 gpurun_out/r2e_pf16.log
env SGG_MPF_BK=32 timeout 200 python tools/mpf_check.py time > gpurun_out/r2e_bk32.log 2>&1; tail -8 gpurun_out/r2e_bk32.log
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
echo "=== bench"; timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err; cut -c1-300 gpurun_out/r2e_bench.json
