#!/bin/bash
mkdir -p gpurun_out
timeout 250 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_tc16<\(int\)1, \(int\)128" -s 1 -c 1 -o gpurun_out/e_prof_lin python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/e_ncu_lin.log 2>&1; echo "rc=$?"
ls -la gpurun_out/e_prof_lin.ncu-rep
