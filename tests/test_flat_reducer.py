"""CPU, world_size = 2 over gloo: FlatGradReducer (flat in-place gradient buffer, fixed bucket order, chunked big
matrices, zero-filled gradients for unused parameters) — the N>1 host logic of SURVEY.md §8e / ADVICE round 1."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgg_b200 import parallel


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.big = torch.nn.Linear(8, 640)         # 5120 weights: larger than the bucket => row chunks
        self.mid = torch.nn.Linear(640, 4)
        self.unused = torch.nn.Linear(4, 4)        # no gradient on rank 1 (rank-dependent control flow)
        self.out = torch.nn.Linear(4, 1)

    def forward(self, x, use_extra):
        h = self.mid(torch.tanh(self.big(x)))
        if use_extra:
            h = h + self.unused(h)
        return self.out(h)


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(12, 8, generator=g), torch.randn(12, 1, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    m = Net()
    red = parallel.FlatGradReducer(m, bucket_bytes=4096, average=True)
    assert len(red.buckets) >= 4 and m.big.weight in red.chunks
    x, y = _data()
    lo, hi = parallel.shard_images(12, rank, world)
    recs = []
    for step in range(3):                          # step 0 learns the arrival order and re-lays the buffer
        red.begin()
        loss = ((m(x[lo:hi], use_extra=(rank == 0)) - y[lo:hi]) ** 2).mean()
        loss.backward()
        red.finish()
        assert all(p.grad is not None and p.grad.data_ptr() == red.view[p].data_ptr() for p in m.parameters())
        recs.append([p.grad.clone() for p in m.parameters()])
    with pytest.raises(RuntimeError):              # a second backward without begin()/finish() is refused, not mis-counted
        red.begin()
        ((m(x[lo:hi], True) - y[lo:hi]) ** 2).mean().backward()
        ((m(x[lo:hi], True) - y[lo:hi]) ** 2).mean().backward()
    if rank == 0:
        torch.save(recs, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_reducer_two_ranks(tmp_path):
    out = str(tmp_path / 'g.pt')
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    recs = torch.load(out)
    x, y = _data()
    # reference: average of the two per-rank gradients computed in one process
    ref = []
    for r in range(2):
        m = Net()
        lo, hi = parallel.shard_images(12, r, 2)
        ((m(x[lo:hi], use_extra=(r == 0)) - y[lo:hi]) ** 2).mean().backward()
        ref.append([p.grad if p.grad is not None else torch.zeros_like(p) for p in m.parameters()])
    for step in range(3):                          # parameters never change (no optimizer) => same gradients every step
        for a, g0, g1 in zip(recs[step], ref[0], ref[1]):
            assert torch.allclose(a, (g0 + g1) / 2, atol=1e-6)


def test_single_process_layout_and_zero_fill():
    m = Net()
    red = parallel.FlatGradReducer(m, bucket_bytes=4096)
    assert red.world == 1 and red.grad_scale == 1.0
    x, y = _data()
    red.begin()
    ((m(x, False) - y) ** 2).mean().backward()
    red.finish()
    assert float(m.unused.weight.grad.abs().max()) == 0.0           # materialised as zeros, not None
    # buckets tile the flat buffer exactly, in order
    assert red.buckets[0].lo == 0 and red.buckets[-1].hi == red.total
    assert all(a.hi == b.lo for a, b in zip(red.buckets, red.buckets[1:]))
    # the learned order puts the last layer's gradients (ready first) at the front
    assert red.offset[m.out.bias] < red.offset[m.big.weight]


def test_shard_images_balanced_and_never_empty():
    assert [parallel.shard_images(5, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 5)]
    with pytest.raises(ValueError):
        parallel.shard_images(3, 0, 4)
