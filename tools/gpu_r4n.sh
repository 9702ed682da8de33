#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/fc6_one.py 2>&1 | tail -1
timeout 300 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:k_lin16p -s 2 -c 1 -f -o gpurun_out/r02f_lin16p python tools/fc6_one.py > gpurun_out/r4n_ncu.log 2>&1; echo rc=$?
ls -la gpurun_out/r02f_lin16p.ncu-rep
