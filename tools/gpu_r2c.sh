#!/bin/bash
mkdir -p gpurun_out
env SGG_TC_TIMING=1 timeout 200 python tools/mpf_check.py time > gpurun_out/r2c_bk64.log 2>&1; cat gpurun_out/r2c_bk64.log
env SGG_MPF_BK=32 SGG_TC_TIMING=1 timeout 200 python tools/mpf_check.py time > gpurun_out/r2c_bk32.log 2>&1; cat gpurun_out/r2c_bk32.log
env SGG_MPF_BK=32 SGG_MPF_PDL=1 timeout 200 python tools/mpf_check.py time > gpurun_out/r2c_bk32_pdl.log 2>&1; cat gpurun_out/r2c_bk32_pdl.log
