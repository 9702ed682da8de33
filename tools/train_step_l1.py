"""PredCls TRAIN step at the L1 boundary (cfg4 per-GPU shape: 32 images x 30 boxes x 300 edges), one process per GPU:
forward + loss (sgg_b200.losses: node CE + 'baseline' edge CE kernels) + CUDA backward + bucketed NCCL gradient all-reduce
(sgg_b200.parallel.GradAllReducer) + fused global-norm clip + SGD(momentum 0.9) sweep (sgg_b200.optim.FusedSGD).  Prints one JSON line on rank 0.

    python tools/train_step.py                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_step.py
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import synth, parallel, losses, optim
from sgg_b200.heads import ImpHeads

world = int(os.environ.get('WORLD_SIZE', 1)); rank = int(os.environ.get('RANK', 0)); lr = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
B = int(os.environ.get('B', 32)); steps = int(os.environ.get('STEPS', 10)); warm = 3
torch.manual_seed(0)
model = ImpHeads().cuda()
if dist is not None:
    for p in model.parameters():
        dist.broadcast(p.data, 0)
opt = optim.FusedSGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
red = parallel.FlatGradReducer(model, bucket_bytes=8 << 20)
batches = []
for i in range(3):
    g = synth.synth_graph(B, 30, 300, 500 + 10 * rank + i)
    N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
    of, ef = synth.synth_l1_feats(N, E, 500 + 10 * rank + i)
    batches.append((torch.from_numpy(of).cuda(), torch.from_numpy(ef).cuda(), torch.from_numpy(g['rel_inds'][:, 1:3].copy()).cuda(),
                    torch.from_numpy(g['gt_classes'][:, 1].copy()).cuda(), torch.from_numpy(g['rel_labels'][:, 3].copy()).cuda()))
loss_hist = []
def step(i):
    of, ef, rel, ocls, rcls = batches[i % 3]
    od, rd = model(of, ef, rel)
    loss = losses.node_losses(od, ocls)['obj_loss'] + losses.edge_losses(rd, rcls, 'baseline')['rel_loss']
    red.begin()
    loss.backward()
    red.finish()
    opt.step(max_norm=5.0, grad_scale=red.grad_scale)   # global-norm clip + 1/world folded into the multi-tensor SGD sweep
    return loss
for i in range(warm):
    step(i)
torch.cuda.synchronize()
if dist is not None: dist.barrier()
t0 = time.perf_counter()
for i in range(steps):
    loss_hist.append(step(i).detach())
torch.cuda.synchronize()
if dist is not None: dist.barrier()
dt = time.perf_counter() - t0
if dist is not None:
    t = torch.tensor([dt], device='cuda', dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t)
if rank == 0:
    ls = [float(x) for x in loss_hist]
    print(json.dumps({'workload': 'PredCls L1 train step (fwd+bwd+grad all-reduce+clip+SGD), %d img/GPU, 30 boxes, 300 edges' % B,
                      'n_gpus': world, 'images_per_s': B * world * steps / dt, 'ms_per_step': dt * 1e3 / steps,
                      'loss_first': ls[0], 'loss_last': ls[-1], 'finite': bool(np.isfinite(ls).all())}))
if dist is not None: dist.destroy_process_group()
if os.environ.get('PROFILE') and rank == 0:
    # GPU-busy fraction of the step: sum of kernel durations (torch profiler, CUPTI) against the wall time of the same steps
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        t0 = time.perf_counter()
        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    busy = sum(e.device_time for e in ev) / 5e3
    print('profile: %.3f ms wall per step (under CUPTI), %.3f ms of kernel time per step, %d kernels per step' % (dt * 1e3 / 5, busy, len(ev) // 5))
    agg = {}
    for e in ev:
        agg.setdefault(e.name[:60], [0, 0.0]); agg[e.name[:60]][0] += 1; agg[e.name[:60]][1] += e.device_time
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
        print('  %-62s %4d %8.1f us/step' % (k, c // 5, t / 5))
