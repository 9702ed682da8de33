"""CPU, world_size = 2 over gloo: image sharding + bucketed gradient all-reduce give the same averaged
gradients as one process on the full batch (the N>1 host logic of SURVEY.md §8e)."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgg_b200 import parallel, synth


def _toy():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 1))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(12, 8, generator=g), torch.randn(12, 1, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    m = _toy()
    red = parallel.GradAllReducer(m, bucket_bytes=256)          # tiny buckets => several collectives
    x, y = _data()
    lo, hi = parallel.shard_images(12, rank, world)
    # per-rank mean loss; equal shard sizes => average of rank means == global mean
    loss = ((m(x[lo:hi]) - y[lo:hi]) ** 2).mean()
    loss.backward()
    red.finish()
    total = parallel.clip_grad_norm(list(m.named_parameters()), 1e9)
    if rank == 0:
        torch.save({'g': [p.grad.clone() for p in m.parameters()], 'norm': total}, out)
    # second step re-uses the reducer (pending counters reset)
    m.zero_grad()
    ((m(x[lo:hi]) - y[lo:hi]) ** 2).mean().backward()
    red.finish()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_grad_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / 'g.pt')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    m = _toy()
    x, y = _data()
    ((m(x) - y) ** 2).mean().backward()
    for a, p in zip(got['g'], m.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-6), (a - p.grad).abs().max()
    ref_norm = torch.sqrt(sum((p.grad ** 2).sum() for p in m.parameters()))
    assert abs(float(got['norm']) - float(ref_norm)) < 1e-5


def test_shard_batch_rebases_image_indices():
    g = synth.synth_graph(4, 5, 6, 3)
    imgs = [torch.zeros(3, 4, 4) for _ in range(4)]
    batch0 = (imgs, None, 0, torch.from_numpy(g['boxes']), torch.from_numpy(g['gt_classes']),
              torch.from_numpy(g['gt_rels']), None, None, ['a', 'b', 'c', 'd'])
    parts = [parallel.shard_batch(batch0, r, 2) for r in range(2)]
    assert [len(p[0]) for p in parts] == [2, 2]
    assert sum(p[3].shape[0] for p in parts) == 20 and sum(p[5].shape[0] for p in parts) == g['gt_rels'].shape[0]
    for p in parts:
        assert set(p[4][:, 0].tolist()) == {0, 1} and set(p[5][:, 0].tolist()) <= {0, 1}
        assert len(p) == len(batch0) and p[-1] in (['a', 'b'], ['c', 'd'])
    assert torch.equal(torch.cat([p[3] for p in parts]), batch0[3])


def test_clip_grad_norm_scales():
    m = _toy()
    x, y = _data()
    ((m(x) - y) ** 2).mean().backward()
    before = torch.sqrt(sum((p.grad ** 2).sum() for p in m.parameters()))
    parallel.clip_grad_norm(list(m.named_parameters()), float(before) / 2)
    after = torch.sqrt(sum((p.grad ** 2).sum() for p in m.parameters()))
    assert abs(float(after) - float(before) / 2) < 1e-4
