"""CPU: the batch wire format (sgg_b200.blob.Blob / vg_collate) against tests/golden/blob.npz, produced by running the
reference's dataloaders/blob.py on the same synthetic entries (tests/golden/make_golden_blob.py); plus the per-rank
slicing used with one process per GPU and the synthetic loader."""
import os
import sys
import numpy as np
import pytest
import torch

from tests import cases
from sgg_b200 import blob as B, parallel

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import make_golden_blob as G  # noqa: E402  (only its CASES / entries / flatten helpers; the reference is not imported)

FX = cases.load('blob')


@pytest.mark.parametrize('name', sorted(G.CASES))
def test_blob_tuple_matches_reference(name):
    mode, is_train, n, td, props, kw = G.CASES[name]
    b = B.vg_collate(G.entries(name), num_gpus=1, is_train=is_train, mode=mode, torch_detector=td, is_cuda=False)
    b = b.scatter()
    assert len(b) == 1
    tup = b[0]
    assert len(tup) == int(FX['%s__tuple_len' % name]) == (9 if is_train else 8)
    got = G.flatten(tup)
    want = {k.split('__', 1)[1]: v for k, v in FX.items() if k.startswith(name + '__') and not k.endswith('tuple_len')}
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k].dtype == want[k].dtype, (k, got[k].dtype, want[k].dtype)
        assert np.array_equal(got[k], want[k]), k


def test_blob_errors_match_reference():
    ds = B.SyntheticVG(num_images=3, with_images=False)
    b = B.Blob(mode='rel', batch_size_per_gpu=2, is_cuda=False, torch_detector=True)
    b.append(ds[0])
    with pytest.raises(ValueError):            # dataloaders/blob.py:147-150
        b.reduce()
    b.append(ds[1]); b.reduce()
    with pytest.raises(ValueError):            # :231-232
        b[1]
    with pytest.raises(AssertionError):
        B.Blob(mode='flickr')
    with pytest.raises(AssertionError):
        B.vg_collate([ds[0]], mode='sgdet')


def test_per_rank_slices_compose_with_shard_batch():
    """num_gpus = 2 (one process per GPU): blob[r] holds rank r's images / boxes / relations with the batch-global image
    index, which parallel.shard_batch-style re-basing turns into a self-contained local batch."""
    ds = B.SyntheticVG(num_images=4, n_box=5, n_rel=3, seed=9, with_images=False)
    full = B.vg_collate([ds[i] for i in range(4)], num_gpus=1, is_train=True, mode='rel', torch_detector=True,
                        is_cuda=False).scatter()[0]
    two = B.vg_collate([ds[i] for i in range(4)], num_gpus=2, is_train=True, mode='rel', torch_detector=True,
                       is_cuda=False).scatter()
    assert len(two) == 2
    for r in range(2):
        imgs, im_sizes, off, boxes, cls, rels, props, _, fns = two[r]
        assert off == 2 * r and len(imgs) == 2 and len(fns) == 2 and im_sizes.shape == (2, 3)
        ref = parallel.shard_batch(full, r, 2)
        assert torch.equal(boxes, ref[3])
        assert torch.equal(cls[:, 1], ref[4][:, 1]) and torch.equal(cls[:, 0] - off, ref[4][:, 0])
        assert torch.equal(rels[:, 1:], ref[5][:, 1:]) and torch.equal(rels[:, 0] - off, ref[5][:, 0])
        assert list(fns) == list(ref[-1])


def test_synthetic_dataset_and_loader_follow_vg_conventions():
    ds = B.SyntheticVG(num_images=8, seed=3, im_hw=(400, 592))
    assert len(ds.ind_to_classes) == 151 and len(ds.ind_to_predicates) == 51
    d = ds[5]
    assert set(d) >= {'img', 'img_size', 'gt_boxes', 'gt_classes', 'gt_relations', 'scale', 'index', 'flipped', 'fn'}
    assert d['img'].shape == (3, d['img_size'][0], d['img_size'][1]) and max(d['img'].shape[1:]) == 592
    assert d['gt_boxes'].dtype == np.float32 and d['gt_classes'].dtype == np.int64 and d['gt_relations'].shape[1] == 3
    assert (d['gt_boxes'][:, 2:] > d['gt_boxes'][:, :2]).all() and d['gt_boxes'].max() <= 1024
    assert d['gt_classes'].min() >= 1 and d['gt_relations'][:, 2].min() >= 1
    assert (d['gt_relations'][:, 0] != d['gt_relations'][:, 1]).all()
    assert np.array_equal(ds[5]['gt_boxes'], d['gt_boxes'])           # deterministic per index
    n = [len(ds[i]['gt_classes']) for i in range(8)]
    assert min(n) >= 2 and max(n) <= 62
    loader = B.synthetic_loader(ds, batch_size=2, is_train=True, is_cuda=False)
    batches = list(loader)
    assert len(batches) == 4
    imgs, im_sizes, off, boxes, cls, rels, props, none, fns = batches[0][0]
    assert isinstance(imgs, list) and len(imgs) == 2 and boxes.dtype == torch.float32 and cls.dtype == torch.int64
    assert boxes.max() <= 592 and rels.shape[1] == 4 and props is None and none is None and len(fns) == 2
