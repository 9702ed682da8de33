#!/bin/bash
# round-2 call B: fused message-passing path — correctness vs SIMT/oracle, timings per variant, full GPU test suite
mkdir -p gpurun_out
run() { echo "=== $1"; shift; env "$@" SGG_TC_TIMING=1 timeout 200 python tools/mpf_check.py all 2>&1 | tail -40; }
run "fused default" SGG_MP_FUSED=1 > gpurun_out/r2b_fused.log 2>&1; cat gpurun_out/r2b_fused.log
env SGG_MP_FUSED=0 timeout 200 python tools/mpf_check.py time > gpurun_out/r2b_old.log 2>&1; cat gpurun_out/r2b_old.log
env SGG_MPF_BK=32 SGG_TC_TIMING=1 timeout 200 python tools/mpf_check.py all > gpurun_out/r2b_bk32.log 2>&1; cat gpurun_out/r2b_bk32.log
env SGG_MPF_PDL=1 timeout 200 python tools/mpf_check.py all > gpurun_out/r2b_pdl.log 2>&1; cat gpurun_out/r2b_pdl.log
echo "=== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
echo "=== bench"; timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err; cut -c1-600 gpurun_out/r2b_bench.json
