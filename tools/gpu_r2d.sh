#!/bin/bash
mkdir -p gpurun_out
env SGG_TC_TIMING=1 timeout 250 python tools/mpf_check.py all > gpurun_out/r2d_bk64.log 2>&1; grep -v "^\[fused.*scale=1 | V.*rd [0-9.e-]*$" gpurun_out/r2d_bk64.log | tail -30
env SGG_MPF_BK=32 SGG_TC_TIMING=1 timeout 200 python tools/mpf_check.py time > gpurun_out/r2d_bk32.log 2>&1; tail -12 gpurun_out/r2d_bk32.log
env SGG_MPF_BK=32 SGG_MPF_PDL=1 timeout 200 python tools/mpf_check.py time > gpurun_out/r2d_bk32_pdl.log 2>&1; tail -8 gpurun_out/r2d_bk32_pdl.log
