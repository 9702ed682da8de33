"""The PredCls TRAIN step of the reference (main.py:100-120) on the full 247.75 M-parameter relation model, from the
``predict`` boundary (RoIAlign-pooled object / union-box features resident on the device):

    res   = model.predict(node_feat, edge_feat, rel_inds, rois, im_sizes)      rel_model_stanford.py:97-107 (train mode:
            union-box geometry branch with batch-statistics BN, fc6/fc7 with dropout, unary, T x message passing, heads)
    loss  = node_losses + edge_losses                                          lib/losses.py:5-74, main.py:105-114
    loss.backward()                                                            main.py:118   (CUDA backward kernels)
    gradient all-reduce over the ranks (991 MB fp32, NEW: the reference is single-GPU, config.py:71)
    grad_clip(model, clip) ; optimizer.step()                                  main.py:119-120 (fused clip + SGD sweep)

The frozen detector (image transform + VGG16 conv stack + RoIAlign) is not part of the timed step: it has no
gradients and no communication; ``bench.py --workload train`` states this in ``config``.
Used by bench.py (BASELINE.json configs[3]: 32 images per GPU, sharded over the ranks) and by tests.
"""
import time

import numpy as np
import torch

from . import losses, optim, parallel, synth


class FakeData(object):
    ind_to_classes = ['__background__'] + ['c%d' % i for i in range(150)]
    ind_to_predicates = ['__background__'] + ['p%d' % i for i in range(50)]


class Conf(object):                      # the fields get_optim reads (config.py defaults)
    l2 = 1e-4
    steps = (15,)
    lr_decay = 0.1


def build_model(device, seed=0, mp_iter=3):
    from .model import RelModelStanford
    torch.manual_seed(seed)              # same seed on every rank: identical replicas without a broadcast
    with torch.device(device):
        m = RelModelStanford(train_data=FakeData(), mode='predcls', mp_iter=mp_iter)
    for _, p in m.detector.named_parameters():
        p.requires_grad = False          # main.py:62-63
    m.train()
    return m


def make_batch(B, boxes, edges, seed, device, pinned=False):
    """Synthetic VG-shaped inputs at the ``predict`` boundary (SURVEY.md §8d generator)."""
    g = synth.synth_graph(B, boxes, edges, seed)
    N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
    rng = np.random.default_rng(seed)
    host = {
        'rois': torch.from_numpy(np.ascontiguousarray(g['rois'])),
        'rel_inds': torch.from_numpy(np.ascontiguousarray(g['rel_inds'])),
        'obj_labels': torch.from_numpy(np.ascontiguousarray(g['gt_classes'][:, 1])),
        'rel_labels': torch.from_numpy(np.ascontiguousarray(g['rel_labels'][:, 3])),
    }
    gen = torch.Generator(device=device).manual_seed(seed)
    dev = {k: v.to(device) for k, v in host.items()}
    # pooled features are post-ReLU VGG activations: relu(N(0,1)); generated on the device (1 GB)
    dev['node_feat'] = torch.relu(torch.randn((N, 512, 7, 7), device=device, generator=gen))
    dev['edge_feat'] = torch.relu(torch.randn((E, 512, 7, 7), device=device, generator=gen))
    del rng
    return dev, N, E


class TrainStep(object):
    def __init__(self, device, B=32, boxes=30, edges=300, mp_iter=3, seed=0, rank=0, lr=1e-3, clip=5.0, comm=True,
                 bucket_bytes=64 << 20, n_batches=2):
        self.device = device
        self.model = build_model(device, seed, mp_iter)
        self.opt, _ = _quiet(lambda: optim.get_optim(self.model, lr, Conf(), -1))
        self.red = parallel.FlatGradReducer(self.model, bucket_bytes=bucket_bytes)
        self.red.comm_enabled = comm
        self.clip = clip
        self.B = B
        self.batches = [make_batch(B, boxes, edges, 4000 + 97 * rank + i, device) for i in range(n_batches)]
        self.N, self.E = self.batches[0][1], self.batches[0][2]
        self.trainable = sum(p.numel() for p in self.model.parameters() if p.requires_grad)
        self.last_loss = None

    def step(self, i, inputs=None):
        d = inputs if inputs is not None else self.batches[i % len(self.batches)][0]
        m = self.model
        od, rd = m.predict(d['node_feat'], d['edge_feat'], d['rel_inds'], d['rois'], None)
        loss = losses.node_losses(od, d['obj_labels'])['obj_loss'] + losses.edge_losses(rd, d['rel_labels'], 'baseline')['rel_loss']
        self.red.begin()
        loss.backward()
        self.red.finish()
        self.opt.step(max_norm=self.clip, grad_scale=self.red.grad_scale)
        self.last_loss = loss.detach()
        return self.last_loss


def _quiet(fn):
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


def allreduce_probe(nbytes, device, reps=5):
    """Stand-alone all-reduce of ``nbytes`` fp32 (the gradient payload): ms per call and NCCL bus bandwidth."""
    import torch.distributed as dist
    world = dist.get_world_size()
    buf = torch.zeros(nbytes // 4, dtype=torch.float32, device=device)
    for _ in range(2):
        dist.all_reduce(buf)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dist.all_reduce(buf)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {'ms': ms, 'bytes': nbytes, 'algbw_gbs': nbytes / ms / 1e6, 'busbw_gbs': nbytes / ms / 1e6 * 2 * (world - 1) / world}


def timed_steps(ts, steps, warmup, barrier):
    for i in range(warmup):
        ts.step(i)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        ts.step(i)
    b.record()
    barrier()
    return a.elapsed_time(b) / steps
