#!/bin/bash
# round-2 call A: host topology, fc6 tile-order experiment (stream-K on/off), memcheck, 1-GPU train step baseline
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | head -30; numactl -H 2>/dev/null; nproc; free -g | head -2;
  for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) class=$(cat $d/class)"; fi; done; } > gpurun_out/r2a_topo.txt 2>&1
echo "=== fc6 default (stream-K)"; timeout 300 python tools/bench_l2.py > gpurun_out/r2a_l2_sk1.json 2> gpurun_out/r2a_l2_sk1.err; echo rc=$?
echo "=== fc6 STREAMK=0"; SGG_TC16_STREAMK=0 timeout 300 python tools/bench_l2.py > gpurun_out/r2a_l2_sk0.json 2> gpurun_out/r2a_l2_sk0.err; echo rc=$?
cat gpurun_out/r2a_l2_sk1.json gpurun_out/r2a_l2_sk0.json | cut -c1-1500
echo "=== train step 1 GPU"; timeout 200 python tools/train_step.py > gpurun_out/r2a_train1.json 2> gpurun_out/r2a_train1.err; echo rc=$?; cat gpurun_out/r2a_train1.json
echo "=== sanitizer memcheck"; timeout 900 bash tools/gpu_sanitize.sh memcheck
echo "=== sanitizer synccheck"; timeout 600 bash tools/gpu_sanitize.sh synccheck
