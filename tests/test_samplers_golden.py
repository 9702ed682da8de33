"""CPU: the training edge samplers (sgg_b200.host) against outputs of the reference's own
lib/proposal_assignments_gtbox.py and lib/rel_assignments.py (tests/golden/samplers.npz, made by
tests/golden/make_golden_samplers.py)."""
import os
import sys
import numpy as np
import pytest
import torch

from tests import cases
from sgg_b200 import host, synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import make_golden_samplers as G  # noqa: E402  (case tables / input builders only)

FX = cases.load('samplers')


@pytest.mark.parametrize('name', sorted(G.GTBOX_CASES))
def test_proposal_assignments_gtbox_equals_reference(name):
    B, nb, ne, seed = G.GTBOX_CASES[name]
    g = synth.synth_graph(B, nb, ne, seed, ragged='ragged' in name)
    rois = torch.from_numpy(g['rois']); cls = torch.from_numpy(g['gt_classes']); rels = torch.from_numpy(g['gt_rels'])
    r, labels, rl = host.proposal_assignments_gtbox(rois, rois[:, 1:], cls, rels, 0, 1024)
    assert r is rois
    assert np.array_equal(labels.numpy(), FX['gtbox_%s_labels' % name])
    assert rl.dtype == torch.int64 and np.array_equal(rl.numpy(), FX['gtbox_%s_rel_labels' % name])


@pytest.mark.parametrize('name', sorted(G.SGDET_CASES))
def test_rel_assignments_equals_reference(name):
    seed, nspg = G.SGDET_CASES[name]
    im, det, lab, gb, gc, gr = G.sgdet_inputs(seed)
    np.random.seed(seed)
    rl = host.rel_assignments(torch.from_numpy(im), torch.from_numpy(det), torch.from_numpy(lab), torch.from_numpy(gb),
                              torch.from_numpy(gc), torch.from_numpy(gr), 0, filter_non_overlap=True,
                              num_sample_per_gt=nspg)
    ref = FX['sgdet_%s' % name]
    got = rl.numpy()
    assert got.dtype == np.int64 and got.shape[1] == 4
    # same numpy RNG stream => same draws; if the call sequence ever diverges this degrades to the semantic checks below
    if got.shape == ref.shape and np.array_equal(got, ref):
        return
    fg_g, fg_r = got[got[:, 3] > 0], ref[ref[:, 3] > 0]
    assert got.shape[0] == ref.shape[0], (got.shape, ref.shape)                     # same budget per image
    assert fg_g.shape[0] == fg_r.shape[0]
    for im_i in np.unique(ref[:, 0]):
        assert (got[:, 0] == im_i).sum() == (ref[:, 0] == im_i).sum()
    assert set(map(tuple, fg_g[:, [0, 3]].tolist())) == set(map(tuple, fg_r[:, [0, 3]].tolist()))
    pytest.fail('rel_assignments matches the reference semantically but not draw-for-draw (RNG call order differs)')


def test_dataset_counts_and_frequency_bias_equal_reference():
    ds = G.FakeCounts()
    for mo in (True, False):
        fg, bg = host.dataset_counts(ds, must_overlap=mo)
        assert np.array_equal(fg, FX['counts_fg_%d' % mo]) and np.array_equal(bg, FX['counts_bg_%d' % mo])
    fg, bg = host.dataset_counts(ds, must_overlap=True)
    fb = host.FrequencyBias(fg, bg)
    assert fb.obj_baseline.weight.shape == FX['freq_weight'].shape
    assert np.abs(fb.obj_baseline.weight.detach().numpy() - FX['freq_weight']).max() <= 1e-6
    out = fb.index_with_labels(torch.from_numpy(FX['freq_labels']))
    assert np.abs(out.detach().numpy() - FX['freq_lookup']).max() <= 1e-6
    fwd = fb(torch.from_numpy(FX['freq_c0']), torch.from_numpy(FX['freq_c1']))
    assert np.abs(fwd.detach().numpy() - FX['freq_forward']).max() <= 1e-5
    assert list(fb.state_dict()) == ['obj_baseline.weight']
