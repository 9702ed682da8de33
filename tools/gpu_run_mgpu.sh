#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "=== N=2 native"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/mgpu2.json 2> gpurun_out/mgpu2.err; echo "rc=$?"; tail -5 gpurun_out/mgpu2.err; cut -c1-700 gpurun_out/mgpu2.json
echo "=== N=2 reference"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/mgpu2_ref.json 2> gpurun_out/mgpu2_ref.err; echo "rc=$?"; cut -c1-200 gpurun_out/mgpu2_ref.json
echo "=== N=1 for comparison"; timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
