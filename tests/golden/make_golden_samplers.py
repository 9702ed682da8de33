#!/usr/bin/env python
"""Golden vectors for the training edge samplers, made by RUNNING THE REFERENCE:
``lib/proposal_assignments_gtbox.py`` (PredCls / SGCls; the cases avoid its CUDA-only ``random_choose`` branch, i.e. no
subsampling, where the result is deterministic) and ``lib/rel_assignments.py`` (SGDet; numpy RNG seeded, its final
``.cuda()`` call patched to a no-op).  Build container only:   python tests/golden/make_golden_samplers.py"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
from sgg_b200 import synth  # noqa: E402
from make_golden import import_reference  # noqa: E402

class FakeCounts(object):
    """Dataset stand-in with the attributes lib/get_dataset_counts.py reads."""
    num_classes, num_predicates = 12, 6

    def __init__(self, seed=21, n_im=9):
        rng = np.random.default_rng(seed)
        self.gt_classes, self.relationships, self.gt_boxes = [], [], []
        for i in range(n_im):
            nb = int(rng.integers(2, 7))
            xy = rng.random((nb, 2)) * 300
            wh = rng.random((nb, 2)) * (40 if i % 3 == 0 else 200) + 5          # every third image: mostly disjoint boxes
            self.gt_boxes.append(np.concatenate((xy, xy + wh), 1).astype(np.float32))
            self.gt_classes.append(rng.integers(1, self.num_classes, nb).astype(np.int64))
            ii, jj = np.nonzero(~np.eye(nb, dtype=bool))
            sel = rng.choice(ii.shape[0], min(3, ii.shape[0]), replace=False)
            self.relationships.append(np.stack((ii[sel], jj[sel], rng.integers(1, self.num_predicates, sel.shape[0])), 1)
                                      .astype(np.int64))

    def __len__(self):
        return len(self.gt_classes)


GTBOX_CASES = {'b3_n6': (3, 6, 10, 5), 'b1_n10': (1, 10, 20, 6), 'b4_ragged': (4, 9, 12, 7)}
SGDET_CASES = {'s1': (11, 1), 's2': (12, 4), 's3': (13, 2)}


def sgdet_inputs(seed):
    """Two images: GT boxes + detections = jittered copies of the GT boxes (some duplicated, some background)."""
    rng = np.random.default_rng(seed)
    g = synth.synth_graph(2, 6, 8, seed)
    gt_boxes = g['boxes']; gt_classes = g['gt_classes']; gt_rels = g['gt_rels']
    det, det_im, det_lab = [], [], []
    for i in range(gt_boxes.shape[0]):
        for _ in range(int(rng.integers(1, 3))):
            det.append(gt_boxes[i] + rng.normal(0, 3.0, 4).astype(np.float32)); det_im.append(gt_classes[i, 0])
            det_lab.append(gt_classes[i, 1] if rng.random() < 0.85 else 0)
    order = np.argsort(np.asarray(det_im), kind='stable')
    return (np.asarray(det_im, np.int64)[order], np.asarray(det, np.float32)[order], np.asarray(det_lab, np.int64)[order],
            gt_boxes, gt_classes, gt_rels)


def main():
    import torch
    import_reference()
    from lib.proposal_assignments_gtbox import proposal_assignments_gtbox
    from lib.rel_assignments import rel_assignments
    out = {}
    for name, (B, nb, ne, seed) in GTBOX_CASES.items():
        g = synth.synth_graph(B, nb, ne, seed, ragged='ragged' in name)
        rois = torch.from_numpy(g['rois']); cls = torch.from_numpy(g['gt_classes']); rels = torch.from_numpy(g['gt_rels'])
        r, labels, rl = proposal_assignments_gtbox(rois, rois[:, 1:], cls, rels, 0, 1024)
        out['gtbox_%s_labels' % name] = labels.numpy(); out['gtbox_%s_rel_labels' % name] = rl.numpy()
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self                 # lib/rel_assignments.py:135 moves the result to the GPU
    try:
        for name, (seed, nspg) in SGDET_CASES.items():
            im, det, lab, gb, gc, gr = sgdet_inputs(seed)
            np.random.seed(seed)
            rl = rel_assignments(torch.from_numpy(im), torch.from_numpy(det), torch.from_numpy(lab), torch.from_numpy(gb),
                                 torch.from_numpy(gc), torch.from_numpy(gr), 0, filter_non_overlap=True,
                                 num_sample_per_gt=nspg)
            out['sgdet_%s' % name] = rl.numpy()
    finally:
        torch.Tensor.cuda = orig_cuda
    # frequency baseline: lib/get_dataset_counts.py + lib/sparse_targets.py (the reference still spells np.float / np.bool)
    if not hasattr(np, 'float'):
        np.float, np.bool = float, bool
    from lib.get_dataset_counts import get_counts
    from lib.sparse_targets import FrequencyBias
    ds = FakeCounts()
    for mo in (True, False):
        fg, bg = get_counts(ds, must_overlap=mo)
        out['counts_fg_%d' % mo] = fg; out['counts_bg_%d' % mo] = bg
    fb = FrequencyBias(ds)
    out['freq_weight'] = fb.obj_baseline.weight.data.numpy()
    lab = torch.from_numpy(np.random.default_rng(3).integers(0, ds.num_classes, (20, 2)))
    out['freq_labels'] = lab.numpy(); out['freq_lookup'] = fb.index_with_labels(lab).detach().numpy()
    c0 = torch.softmax(torch.from_numpy(np.random.default_rng(4).standard_normal((5, ds.num_classes)).astype(np.float32)), 1)
    c1 = torch.softmax(torch.from_numpy(np.random.default_rng(5).standard_normal((5, ds.num_classes)).astype(np.float32)), 1)
    out['freq_c0'] = c0.numpy(); out['freq_c1'] = c1.numpy(); out['freq_forward'] = fb(c0, c1).detach().numpy()
    np.savez_compressed(os.path.join(HERE, 'samplers.npz'), **out)
    print('wrote samplers.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
