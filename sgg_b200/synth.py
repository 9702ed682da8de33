"""Synthetic Visual-Genome-shaped inputs and weights.

There is no dataset / checkpoint access, so every test, fixture and benchmark
uses inputs generated here from a numpy ``default_rng`` seed.  The same
generator is used by the golden-fixture script (which feeds the reference), the
oracle tests and ``bench.py``, so "identical inputs" is by construction.

Shapes follow the reference: boxes are pixel xyxy in the 592-px frame
(config.py:32 IM_SCALE), classes 1..150, predicates 1..50, edges are ordered
pairs (i != j) of same-image objects sorted by (img, subj, obj) exactly as
lib/proposal_assignments_gtbox.py:74-77 leaves them.
"""
import hashlib
import numpy as np

IM_SCALE = 592
NUM_CLASSES = 151
NUM_RELS = 51


def synth_graph(B, n_box, n_edge, seed, ragged=False, all_pairs=False):
    """Returns a dict with
    boxes [N,4] f32, gt_classes [N,2] i64 (img, cls), im_inds [N] i64,
    rel_inds [E,3] i64 (img, subj_global, obj_global) sorted,
    gt_rels [R,4] i64 (img, subj_local, obj_local, pred), rel_labels [E,4] i64.
    """
    rng = np.random.default_rng(seed)
    boxes, classes, rel_inds, gt_rels, rel_labels = [], [], [], [], []
    base = 0
    for b in range(B):
        nb = n_box
        if ragged:
            nb = int(np.clip(np.rint(rng.normal(n_box, 0.2 * n_box)), 2, 62))
        xy = rng.random((nb, 2), dtype=np.float32) * np.float32(0.6 * IM_SCALE)
        wh = rng.random((nb, 2), dtype=np.float32) * np.float32(0.4 * IM_SCALE - 16) + np.float32(16)
        boxes.append(np.concatenate((xy, xy + wh), 1).astype(np.float32))
        cls = rng.integers(1, NUM_CLASSES, nb)
        classes.append(np.stack((np.full(nb, b), cls), 1))
        ii, jj = np.nonzero(~np.eye(nb, dtype=bool))
        ncand = ii.shape[0]
        if all_pairs or n_edge >= ncand:
            sel = np.arange(ncand)
        else:
            sel = np.sort(rng.choice(ncand, n_edge, replace=False))
        si, oi = ii[sel], jj[sel]
        pred = np.zeros(sel.shape[0], np.int64)
        nfg = min(5, sel.shape[0])
        fg = rng.choice(sel.shape[0], nfg, replace=False)
        pred[fg] = rng.integers(1, NUM_RELS, nfg)
        rel_inds.append(np.stack((np.full_like(si, b), si + base, oi + base), 1))
        rel_labels.append(np.stack((np.full_like(si, b), si + base, oi + base, pred), 1))
        gt_rels.append(np.stack((np.full(nfg, b), si[fg], oi[fg], pred[fg]), 1))
        base += nb
    g = dict(boxes=np.concatenate(boxes), gt_classes=np.concatenate(classes).astype(np.int64),
             rel_inds=np.concatenate(rel_inds).astype(np.int64),
             gt_rels=np.concatenate(gt_rels).astype(np.int64),
             rel_labels=np.concatenate(rel_labels).astype(np.int64))
    g['im_inds'] = g['gt_classes'][:, 0].copy()
    g['rois'] = np.concatenate((g['im_inds'][:, None].astype(np.float32), g['boxes']), 1)
    return g


def synth_l1_feats(N, E, seed, D=4096):
    """Precomputed-feature inputs at the L1 boundary: node head ends in ReLU
    (rel_model_base.py:111), edge head ends linear (:110)."""
    rng = np.random.default_rng(seed + 7919)
    obj = np.maximum(rng.standard_normal((N, D), dtype=np.float32), 0)
    edge = rng.standard_normal((E, D), dtype=np.float32)
    return obj, edge


def synth_l0_states(N, E, seed, H=512):
    """message_pass inputs: obj_rep (obj_unary output, linear) and rel_rep (post-ReLU)."""
    rng = np.random.default_rng(seed + 104729)
    obj = rng.standard_normal((N, H), dtype=np.float32) * np.float32(0.5)
    rel = np.maximum(rng.standard_normal((E, H), dtype=np.float32) * np.float32(0.5), 0)
    return obj, rel


def synth_pooled(N, E, seed, C=512, P=7):
    """L2 inputs: RoIAlign outputs of a post-ReLU feature map are non-negative."""
    rng = np.random.default_rng(seed + 15485863)
    node = np.maximum(rng.standard_normal((N, C, P, P), dtype=np.float32), 0)
    edge = np.maximum(rng.standard_normal((E, C, P, P), dtype=np.float32), 0)
    return node, edge


def synth_fmap(B, seed, C=512, S=38):
    rng = np.random.default_rng(seed + 32452843)
    return np.maximum(rng.standard_normal((B, C, S, S), dtype=np.float32), 0)


def _u(rng, shape, bound):
    return ((rng.random(shape, dtype=np.float32) * 2 - 1) * np.float32(bound)).astype(np.float32)


def synth_params(seed, H=512, D=4096, scale=1.0, level='l1', C=512, P=7):
    """State-dict-keyed weights (names/shapes of SURVEY.md §8a), U(-s/sqrt(fan_in), s/sqrt(fan_in))
    like torch's default Linear/GRUCell init, ``scale`` > 1 emulates trained logit magnitudes.

    level: 'l0' message-passing only; 'l1' + unary/heads; 'l2' + union_boxes.conv and roi_fmap*.
    """
    rng = np.random.default_rng(seed + 611953)
    p = {}
    s = float(scale)
    for gname in ('edge_gru', 'node_gru'):
        b = s / np.sqrt(H)
        p[gname + '.weight_ih'] = _u(rng, (3 * H, H), b)
        p[gname + '.weight_hh'] = _u(rng, (3 * H, H), b)
        p[gname + '.bias_ih'] = _u(rng, (3 * H,), b)
        p[gname + '.bias_hh'] = _u(rng, (3 * H,), b)
    for k in ('sub_vert', 'obj_vert', 'out_edge', 'in_edge'):
        b = s / np.sqrt(2 * H)
        p[k + '_w_fc.0.weight'] = _u(rng, (1, 2 * H), b)
        p[k + '_w_fc.0.bias'] = _u(rng, (1,), b)
    if level in ('l1', 'l2'):
        for k, (o, i) in (('obj_unary', (H, D)), ('edge_unary', (H, D)),
                          ('obj_fc', (NUM_CLASSES, H)), ('rel_fc', (NUM_RELS, H))):
            b = s / np.sqrt(i)
            p[k + '.weight'] = _u(rng, (o, i), b)
            p[k + '.bias'] = _u(rng, (o,), b)
    if level == 'l2':
        half = C // 2
        p['union_boxes.conv.0.weight'] = _u(rng, (half, 2, 7, 7), s / np.sqrt(2 * 49))
        p['union_boxes.conv.0.bias'] = _u(rng, (half,), s / np.sqrt(2 * 49))
        p['union_boxes.conv.4.weight'] = _u(rng, (C, half, 3, 3), s / np.sqrt(half * 9))
        p['union_boxes.conv.4.bias'] = _u(rng, (C,), s / np.sqrt(half * 9))
        for idx, ch in (('2', half), ('6', C)):
            p['union_boxes.conv.%s.weight' % idx] = (rng.random(ch, dtype=np.float32) + np.float32(0.5))
            p['union_boxes.conv.%s.bias' % idx] = _u(rng, (ch,), 0.2)
            p['union_boxes.conv.%s.running_mean' % idx] = _u(rng, (ch,), 0.2)
            p['union_boxes.conv.%s.running_var' % idx] = (rng.random(ch, dtype=np.float32) + np.float32(0.5))
        fin = C * P * P
        for pre in ('roi_fmap.1.', 'roi_fmap_obj.'):
            p[pre + '0.weight'] = _u(rng, (D, fin), s / np.sqrt(fin))
            p[pre + '0.bias'] = _u(rng, (D,), s / np.sqrt(fin))
            p[pre + '3.weight'] = _u(rng, (D, D), s / np.sqrt(D))
            p[pre + '3.bias'] = _u(rng, (D,), s / np.sqrt(D))
    return p


def digest(*arrays):
    """Short content hash used by fixtures to detect generator drift."""
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def synth_backbone_params(model_state_shapes, seed):
    """He-scaled weights for the frozen detector backbone (VGG16 conv stack): keeps activations O(1) so the
    L3 fixtures exercise a realistic feature map.  ``model_state_shapes``: {key: shape} of detector.backbone.*"""
    rng = np.random.default_rng(seed + 2750159)
    p = {}
    for k in sorted(model_state_shapes):
        shp = tuple(model_state_shapes[k])
        if k.endswith('.weight'):
            fan_in = int(np.prod(shp[1:]))
            p[k] = (rng.standard_normal(shp, dtype=np.float32) * np.float32(np.sqrt(2.0 / fan_in))).astype(np.float32)
        else:
            p[k] = _u(rng, shp, 0.05)
    return p


def synth_images(sizes, seed):
    rng = np.random.default_rng(seed + 86028121)
    return [rng.random((3, h, w), dtype=np.float32) for (h, w) in sizes]


# ---- training tail (losses / clipping / SGD): inputs of tests/golden/train_tail.npz --------------------------
def synth_logits(M, C, seed, fg=0.15, scale=3.0):
    """Logits [M,C] ~ scale * N(0,1) and int64 labels with a fraction ``fg`` of foreground rows (label in 1..C-1),
    the rest background (0) — the shape of rel_labels[:, -1] (proposal_assignments_gtbox.py:70-77)."""
    rng = np.random.default_rng(9000 + seed)
    logits = (scale * rng.standard_normal((M, C))).astype(np.float32)
    labels = np.zeros(M, np.int64)
    if M > 0 and fg > 0:
        is_fg = rng.random(M) < fg
        if fg >= 1.0:
            is_fg[:] = True
        labels[is_fg] = rng.integers(1, C, int(is_fg.sum()))
    return logits, labels


def explicit_idx(labels, seed):
    """Explicit idx_fg / idx_bg (lib/losses.py:27-31 lets the caller pass them): a subset of the FG rows and a
    subset of the BG rows, so some rows are in neither set and keep weight 1."""
    rng = np.random.default_rng(9100 + seed)
    fg = np.nonzero(labels > 0)[0]
    bg = np.nonzero(labels == 0)[0]
    return np.sort(rng.choice(fg, max(1, len(fg) // 2), replace=False)), \
        np.sort(rng.choice(bg, max(1, (2 * len(bg)) // 3), replace=False))


def loss_cases():
    return {
        'node_cfg2': dict(kind='node', M=240, C=151, seed=1, fg=1.0, w=(1, 1, 1)),
        'baseline_cfg2': dict(kind='baseline', M=2400, C=51, seed=2, fg=0.02, w=(1, 1, 1)),
        'baseline_gamma': dict(kind='baseline', M=300, C=51, seed=3, fg=0.1, w=(1, 1, 0.5)),
        'dnorm_cfg2': dict(kind='dnorm', M=2400, C=51, seed=4, fg=0.02, w=(1, 1, 1)),
        'dnorm_weights': dict(kind='dnorm', M=777, C=51, seed=5, fg=0.2, w=(0.5, 2.0, 1.5)),
        'dnorm_no_fg': dict(kind='dnorm', M=64, C=51, seed=6, fg=0.0, w=(1, 1, 1)),       # weights stay 1 (:51,:57)
        'dnorm_no_bg': dict(kind='dnorm', M=64, C=51, seed=7, fg=1.0, w=(1, 1, 1)),
        'fgbg_cfg2': dict(kind='dnorm-fgbg', M=2400, C=51, seed=8, fg=0.02, w=(1, 1, 1)),
        'fgbg_weights': dict(kind='dnorm-fgbg', M=1001, C=51, seed=9, fg=0.3, w=(2.0, 0.25, 3.0)),
        'fgbg_no_fg': dict(kind='dnorm-fgbg', M=33, C=51, seed=10, fg=0.0, w=(1, 1, 1)),
        'dnorm_explicit': dict(kind='dnorm', M=500, C=51, seed=11, fg=0.2, w=(1, 1, 1), explicit_idx=True),
        'one_row': dict(kind='dnorm', M=1, C=51, seed=12, fg=1.0, w=(1, 1, 1)),
    }


def synth_train_tail(seed=5, steps=3):
    """A small parameter set with the reference's naming (two ``roi_fmap*`` tensors fall in the lr/10 group of
    get_optim, lib/pytorch_misc.py:135-142), sizes that exercise the vector path, ragged tails, a tiny tensor and a
    tensor that gets no gradient on some steps; gradients are large enough that clipping at 5.0 engages on step 0
    and small enough that it does not on step 2."""
    rng = np.random.default_rng(9200 + seed)
    shapes = {'roi_fmap.1.0.weight': (8, 1024), 'roi_fmap_obj.0.bias': (4099,), 'edge_gru.weight_ih': (8, 512),
              'obj_fc.weight': (9, 512), 'obj_fc.bias': (151,), 'sub_vert_w_fc.0.bias': (1,),
              'union_boxes.conv.0.weight': (16, 2, 7, 7), 'rel_fc.weight': (3, 512)}
    params = {k: (0.1 * rng.standard_normal(s)).astype(np.float32) for k, s in shapes.items()}
    gscale = [0.05, 0.01, 0.0005][:steps] + [0.0005] * max(0, steps - 3)
    grads = []
    for t in range(steps):
        g = {k: (gscale[t] * rng.standard_normal(s)).astype(np.float32) for k, s in shapes.items()}
        if t == 1:
            g['rel_fc.weight'] = None          # no gradient this step: torch skips the tensor entirely
        grads.append(g)
    return dict(params=params, grads=grads, lr=0.12, l2=1e-4, clip=5.0, steps=steps)
