#!/bin/bash
# bench at N ranks the way the driver launches it
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  SGG_BENCH_WATCHDOG=500 timeout 600 python bench.py --gpus 1 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
else
  SGG_BENCH_WATCHDOG=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
fi
echo rc=$?; grep -E "^\[bench" gpurun_out/scale_n$N.err | tail -8; tail -c 600 gpurun_out/scale_n$N.json
