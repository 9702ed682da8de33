#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 2 2 1; do
echo "=== PDL=$pdl"; SGG_MPF_PDL=$pdl timeout 200 python tools/mpf_check.py time 2>&1 | grep -E "l1 cuda-graph|message_pass T=3|B=8 N=240"
done
echo "=== PDL=2 check"; SGG_MPF_PDL=2 timeout 250 python tools/mpf_check.py check 2>&1 | tail -8
echo "=== PDL=2 pytest parity x2"; for i in 1 2; do SGG_MPF_PDL=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py tests/test_gpu_grad.py -m gpu -x -q 2>&1 | tail -1; done
