"""PredCls TRAIN step (sgg_b200/trainstep.py: full 247.75 M-parameter model from the predict boundary, fwd + losses +
CUDA backward + flat-buffer NCCL gradient all-reduce + fused clip + SGD), one process per GPU.  One JSON line on rank 0.

    python tools/train_step.py                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_step.py
Env: B (images per GPU, 32), STEPS (8), COMM=0 (skip the collectives), SYNC=1 (synchronise after every stage: debugging).
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import trainstep

world = int(os.environ.get('WORLD_SIZE', 1)); rank = int(os.environ.get('RANK', 0)); lr = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)
B = int(os.environ.get('B', 32)); steps = int(os.environ.get('STEPS', 8))


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


ts = trainstep.TrainStep(dev, B=B, rank=rank, comm=os.environ.get('COMM', '1') != '0')
print('rank %d: model built, N=%d E=%d' % (rank, ts.N, ts.E), file=sys.stderr, flush=True)
first = float(ts.step(0)); torch.cuda.synchronize()
print('rank %d: first step ok, loss %.4f' % (rank, first), file=sys.stderr, flush=True)
ms = trainstep.timed_steps(ts, steps, 2, barrier)
if dist is not None:
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
if rank == 0:
    print(json.dumps({'workload': 'PredCls train step (predict boundary), %d img/GPU' % B, 'n_gpus': world,
                      'images_per_s': B * world / (ms * 1e-3), 'ms_per_step': ms, 'loss_first': first,
                      'loss_last': float(ts.last_loss), 'buckets': len(ts.red.buckets),
                      'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}))
if dist is not None:
    dist.destroy_process_group()
