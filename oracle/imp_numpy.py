"""CPU oracle (numpy) for the IMP relation-model hot path of bknyaz/sgg.

TEST INFRASTRUCTURE ONLY.  This module is the *checker* for the CUDA path in
``sgg_b200``; it is never imported by the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.

Every function restates, from the reference's mathematics, one function of the
hot path and cites the reference file:line it follows (paths relative to the
reference repo root).  Third-party arithmetic the reference delegates to
(torch ``nn.GRUCell`` / ``nn.Linear`` / ``BatchNorm2d`` / ``Conv2d`` /
``MaxPool2d`` and torchvision ``roi_align``; both unpinned by the reference,
torch 2.11.0 / torchvision 0.26.0 in this image) is restated from its published
definition.

Pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md §4, §8c), so the oracle is pinned against outputs of the reference
itself, run in the build container by ``tests/golden/make_golden.py`` and stored
under ``tests/golden/*.npz`` (see ``tests/test_oracle_golden.py``).

All arithmetic is float32 unless ``dtype=np.float64`` is requested (used to
measure the fp32 noise floor).
"""
import numpy as np

H_DEFAULT = 512


# --------------------------------------------------------------------------- #
# elementary ops (third-party arithmetic restated)
# --------------------------------------------------------------------------- #
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def linear(x, w, b=None):
    """torch.nn.Linear: y = x @ w.T + b   (w is [out, in])."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRUCell (sgg_models/rel_model_stanford.py:36-37, used :71,72,83,92).

    r = s(W_ir x + b_ir + W_hr h + b_hr); z = s(W_iz x + b_iz + W_hz h + b_hz)
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1 - z) * n + z * h
    weight_ih / weight_hh are [3H, H] in (r, z, n) row-block order.
    """
    Hd = h.shape[1]
    gi = x @ w_ih.T + b_ih
    gh = h @ w_hh.T + b_hh
    r = sigmoid(gi[:, :Hd] + gh[:, :Hd])
    z = sigmoid(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
    n = np.tanh(gi[:, 2 * Hd:] + r * gh[:, 2 * Hd:])
    return ((1.0 - z) * n + z * h).astype(x.dtype)


def _gate(w, b, vert, edge):
    """nn.Sequential(Linear(2H,1), Sigmoid) on cat(vert, edge)
    (rel_model_stanford.py:41-45, used :78-81,86-89).  w:[1,2H], b:[1]."""
    Hd = vert.shape[1]
    s = vert @ w[0, :Hd] + edge @ w[0, Hd:] + b[0]
    return sigmoid(s)[:, None].astype(vert.dtype)


# --------------------------------------------------------------------------- #
# a1: message_pass
# --------------------------------------------------------------------------- #
def message_pass(rel_rep, obj_rep, rel_inds, p, mp_iter=3, return_all=False):
    """RelModelStanford.message_pass (rel_model_stanford.py:48-94).

    rel_rep [E,H], obj_rep [N,H], rel_inds [E,2] (global subject, object ids).
    ``p`` maps the reference's state-dict keys to numpy arrays.
    Both updates of iteration i read only iteration-i states (Jacobi), and the
    vertex context uses the OLD edge state (:86-89).  The reference's dense
    [N,E] incidence products (:58-66, :91) are restated as index scatter-adds.
    """
    dt = rel_rep.dtype
    N, E = obj_rep.shape[0], rel_rep.shape[0]
    Hd = p['edge_gru.weight_hh'].shape[1]
    sub, ob = rel_inds[:, 0].astype(np.int64), rel_inds[:, 1].astype(np.int64)
    eg = [p['edge_gru.' + k].astype(dt) for k in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')]
    ng = [p['node_gru.' + k].astype(dt) for k in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')]
    gw = {k: (p[k + '_w_fc.0.weight'].astype(dt), p[k + '_w_fc.0.bias'].astype(dt))
          for k in ('sub_vert', 'obj_vert', 'out_edge', 'in_edge')}

    vert = [gru_cell(obj_rep, np.zeros((N, Hd), dt), *ng)]      # :68-71
    edge = [gru_cell(rel_rep, np.zeros((E, Hd), dt), *eg)]      # :72
    for i in range(mp_iter):                                    # :74
        sv, ov = vert[i][sub], vert[i][ob]                      # :76-77
        w_sub = _gate(*gw['sub_vert'], sv, edge[i]) * sv        # :78-79
        w_obj = _gate(*gw['obj_vert'], ov, edge[i]) * ov        # :80-81
        edge.append(gru_cell(w_sub + w_obj, edge[i], *eg))      # :83
        pre_out = _gate(*gw['out_edge'], sv, edge[i]) * edge[i]  # :86-87
        pre_in = _gate(*gw['in_edge'], ov, edge[i]) * edge[i]    # :88-89
        ctx = np.zeros((N, Hd), dt)
        np.add.at(ctx, sub, pre_out)                            # objs_to_outrels @ pre_out :91
        np.add.at(ctx, ob, pre_in)                              # objs_to_inrels @ pre_in  :91
        vert.append(gru_cell(ctx, vert[i], *ng))                # :92
    if return_all:
        return vert, edge
    return vert[-1], edge[-1]


# --------------------------------------------------------------------------- #
# L1: precomputed 4096-d features -> dists  (rel_model_stanford.py:103-107 minus roi_fmap*)
# --------------------------------------------------------------------------- #
def l1_forward(obj_feat, edge_feat, rel_inds, p, mp_iter=3):
    """obj_unary -> edge_unary+ReLU -> message_pass -> obj_fc / rel_fc.
    obj_feat [N,4096], edge_feat [E,4096], rel_inds [E,2]."""
    dt = obj_feat.dtype
    nf = linear(obj_feat, p['obj_unary.weight'].astype(dt), p['obj_unary.bias'].astype(dt))
    ef = np.maximum(linear(edge_feat, p['edge_unary.weight'].astype(dt), p['edge_unary.bias'].astype(dt)), 0)
    v, e = message_pass(ef, nf, rel_inds, p, mp_iter)
    return (linear(v, p['obj_fc.weight'].astype(dt), p['obj_fc.bias'].astype(dt)),
            linear(e, p['rel_fc.weight'].astype(dt), p['rel_fc.bias'].astype(dt)))


# --------------------------------------------------------------------------- #
# a8: draw_union_boxes (the reference's only native code, Cython)
# --------------------------------------------------------------------------- #
def draw_union_boxes(box_pairs, pooling_size=27):
    """lib/draw_rectangles/draw_rectangles.pyx:12-67.

    box_pairs [E,8] f32 = (x1,y1,x2,y2) of subject then object.  Output
    [E,2,P,P]: anti-aliased occupancy of each box in union-box coordinates,
    cell (j,k) = clamp01(k+1-x1)*clamp01(x2-k) * clamp01(j+1-y1)*clamp01(y2-j).
    float32 arithmetic in the same operation order as the pyx (:45-66).
    """
    bp = np.asarray(box_pairs, np.float32)
    E = bp.shape[0]
    P = np.float32(pooling_size)
    x1u = np.minimum(bp[:, 0], bp[:, 4]); y1u = np.minimum(bp[:, 1], bp[:, 5])
    x2u = np.maximum(bp[:, 2], bp[:, 6]); y2u = np.maximum(bp[:, 3], bp[:, 7])
    w = (x2u - x1u).astype(np.float32); h = (y2u - y1u).astype(np.float32)
    out = np.zeros((E, 2, pooling_size, pooling_size), np.float32)
    grid = np.arange(pooling_size, dtype=np.float32)
    c01 = lambda a: np.minimum(np.maximum(a, np.float32(0)), np.float32(1))
    with np.errstate(divide='ignore', invalid='ignore'):
        for i in range(2):
            x1 = ((bp[:, 0 + 4 * i] - x1u) * P / w).astype(np.float32)
            y1 = ((bp[:, 1 + 4 * i] - y1u) * P / h).astype(np.float32)
            x2 = ((bp[:, 2 + 4 * i] - x1u) * P / w).astype(np.float32)
            y2 = ((bp[:, 3 + 4 * i] - y1u) * P / h).astype(np.float32)
            yc = c01(grid[None] + np.float32(1) - y1[:, None]) * c01(y2[:, None] - grid[None])   # [E,P]
            xc = c01(grid[None] + np.float32(1) - x1[:, None]) * c01(x2[:, None] - grid[None])
            out[:, i] = (xc[:, None, :] * yc[:, :, None]).astype(np.float32)
    return out


# --------------------------------------------------------------------------- #
# a7: UnionBoxesAndFeats geometry branch
# --------------------------------------------------------------------------- #
def _conv2d(x, w, b, stride, pad):
    """torch.nn.Conv2d (cross-correlation), NCHW, square stride/pad."""
    n, c, hh, ww = x.shape
    co, ci, kh, kw = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    ho = (hh + 2 * pad - kh) // stride + 1
    wo = (ww + 2 * pad - kw) // stride + 1
    out = np.empty((n, co, ho, wo), x.dtype)
    w2 = w.reshape(co, -1)
    for i in range(ho):
        for j in range(wo):
            patch = xp[:, :, i * stride:i * stride + kh, j * stride:j * stride + kw].reshape(n, -1)
            out[:, :, i, j] = patch @ w2.T + b
    return out


def _maxpool2d(x, k, stride, pad):
    n, c, hh, ww = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)), constant_values=-np.inf)
    ho = (hh + 2 * pad - k) // stride + 1
    wo = (ww + 2 * pad - k) // stride + 1
    out = np.empty((n, c, ho, wo), x.dtype)
    for i in range(ho):
        for j in range(wo):
            out[:, :, i, j] = xp[:, :, i * stride:i * stride + k, j * stride:j * stride + k].max((2, 3))
    return out


def _batchnorm2d(x, g, b, rm, rv, training, eps=1e-5):
    if training:
        m = x.mean((0, 2, 3)); v = x.var((0, 2, 3))            # biased batch stats
    else:
        m, v = rm, rv
    sh = (1, -1, 1, 1)
    return ((x - m.reshape(sh)) / np.sqrt(v.reshape(sh) + eps) * g.reshape(sh) + b.reshape(sh)).astype(x.dtype)


def union_geom(rects, p, training=False, prefix='union_boxes.conv.'):
    """UnionBoxesAndFeats.conv (lib/get_union_boxes.py:51-59) applied to
    ``rects = draw_union_boxes(...) - 0.5`` (:67).

    Both convs run with stride 16 because the ``conv_layer`` lambda captures the
    constructor's ``stride`` (:40-43): [E,2,27,27] -> conv7/s16/p3 -> [E,256,2,2]
    -> ReLU -> BN -> maxpool(3,2,1) -> [E,256,1,1] -> conv3/s16/p1 -> [E,512,1,1]
    -> ReLU -> BN.  Returns [E,512,1,1]."""
    dt = rects.dtype
    g = lambda k: p[prefix + k].astype(dt)
    x = _conv2d(rects, g('0.weight'), g('0.bias'), 16, 3)
    x = np.maximum(x, 0)
    x = _batchnorm2d(x, g('2.weight'), g('2.bias'), g('2.running_mean'), g('2.running_var'), training)
    x = _maxpool2d(x, 3, 2, 1)
    x = _conv2d(x, g('4.weight'), g('4.bias'), 16, 1)
    x = np.maximum(x, 0)
    x = _batchnorm2d(x, g('6.weight'), g('6.bias'), g('6.running_mean'), g('6.running_var'), training)
    return x


def union_boxes_and_feats(union_pools, rois, union_inds, p, training=False):
    """UnionBoxesAndFeats.forward, edge_model='motifs' (lib/get_union_boxes.py:63-101):
    union_pools [E,C,7,7] + conv(draw_union_boxes(pair_rois, 27) - 0.5) (broadcast [E,C,1,1])."""
    pair = np.concatenate((rois[:, 1:][union_inds[:, 0]], rois[:, 1:][union_inds[:, 1]]), 1)
    rects = (draw_union_boxes(pair, 27) - np.float32(0.5)).astype(union_pools.dtype)
    return union_pools + union_geom(rects, p, training)


# --------------------------------------------------------------------------- #
# a9: node_edge_features = union boxes + torchvision roi_align
# --------------------------------------------------------------------------- #
def union_rois(rois, union_inds):
    """rel_model_base.py:248-250: (img, min x1y1, max x2y2) of the pair."""
    a, b = rois[union_inds[:, 0]], rois[union_inds[:, 1]]
    return np.concatenate((a[:, :1], np.minimum(a[:, 1:3], b[:, 1:3]), np.maximum(a[:, 3:5], b[:, 3:5])), 1)


def roi_align(fmap, rois, out_size=7, spatial_scale=1.0 / 16, sampling_ratio=2):
    """torchvision.ops.roi_align, aligned=False (called through MultiScaleRoIAlign,
    rel_model_base.py:97-99,258-259).  fmap [B,C,Hf,Wf], rois [R,5]=(img,x1,y1,x2,y2).

    roi_w = max(x2*s - x1*s, 1); bin = roi/out; sampling grid sr x sr per bin at
    y = y1*s + ph*bin_h + (iy+.5)*bin_h/sr; bilinear with the torchvision edge
    rules (sample outside [-1, size] contributes 0; y<=0 -> 0; clamp at size-1)."""
    B, C, Hf, Wf = fmap.shape
    R = rois.shape[0]
    out = np.zeros((R, C, out_size, out_size), fmap.dtype)
    f32 = np.float32
    for r in range(R):
        b = int(rois[r, 0])
        x1, y1, x2, y2 = [f32(v) * f32(spatial_scale) for v in rois[r, 1:5]]
        rw = max(x2 - x1, f32(1.0)); rh = max(y2 - y1, f32(1.0))
        bw = f32(rw) / f32(out_size); bh = f32(rh) / f32(out_size)
        sr = sampling_ratio
        for ph in range(out_size):
            for pw in range(out_size):
                acc = np.zeros(C, np.float64 if fmap.dtype == np.float64 else np.float32)
                for iy in range(sr):
                    y = y1 + f32(ph) * bh + (f32(iy) + f32(0.5)) * bh / f32(sr)
                    for ix in range(sr):
                        x = x1 + f32(pw) * bw + (f32(ix) + f32(0.5)) * bw / f32(sr)
                        if y < -1.0 or y > Hf or x < -1.0 or x > Wf:
                            continue
                        yy = max(y, f32(0)); xx = max(x, f32(0))
                        yl = int(yy); xl = int(xx)
                        if yl >= Hf - 1:
                            yh = yl = Hf - 1; yy = f32(yl)
                        else:
                            yh = yl + 1
                        if xl >= Wf - 1:
                            xh = xl = Wf - 1; xx = f32(xl)
                        else:
                            xh = xl + 1
                        ly = f32(yy - yl); lx = f32(xx - xl)
                        hy = f32(1.0) - ly; hx = f32(1.0) - lx
                        acc += (hy * hx * fmap[b, :, yl, xl] + hy * lx * fmap[b, :, yl, xh]
                                + ly * hx * fmap[b, :, yh, xl] + ly * lx * fmap[b, :, yh, xh])
                out[r, :, ph, pw] = acc / f32(sr * sr)
    return out


def node_edge_features(fmap, rois, union_inds):
    """RelModelBase.node_edge_features (rel_model_base.py:245-260), vgg16 branch.
    MultiScaleRoIAlign with one level infers scale 2^round(log2(38/592)) = 1/16."""
    return roi_align(fmap, rois), roi_align(fmap, union_rois(rois, union_inds))


# --------------------------------------------------------------------------- #
# a6 / a4: feature heads and predict
# --------------------------------------------------------------------------- #
def roi_fmap_edge(x, p):
    """roi_fmap = Flatten, Linear(25088,4096), ReLU, Dropout, Linear(4096,4096)
    (rel_model_base.py:92,110; eval mode: Dropout is identity).  No final ReLU."""
    x = x.reshape(x.shape[0], -1)
    x = np.maximum(linear(x, p['roi_fmap.1.0.weight'], p['roi_fmap.1.0.bias']), 0)
    return linear(x, p['roi_fmap.1.3.weight'], p['roi_fmap.1.3.bias'])


def roi_fmap_node(x, p):
    """roi_fmap_obj = Linear, ReLU, Dropout, Linear, ReLU, Dropout (rel_model_base.py:111)."""
    x = x.reshape(x.shape[0], -1)
    x = np.maximum(linear(x, p['roi_fmap_obj.0.weight'], p['roi_fmap_obj.0.bias']), 0)
    return np.maximum(linear(x, p['roi_fmap_obj.3.weight'], p['roi_fmap_obj.3.bias']), 0)


def predict(node_feat, edge_feat, rel_inds, rois, p, mp_iter=3, training_bn=False):
    """RelModelStanford.predict (rel_model_stanford.py:97-107), eval-mode dropout.
    node_feat [N,C,7,7], edge_feat [E,C,7,7], rel_inds [E,3]=(img,subj,obj), rois [N,5]."""
    ef = union_boxes_and_feats(edge_feat, rois, rel_inds[:, 1:], p, training_bn)     # :100-101
    nf4096 = roi_fmap_node(node_feat, p)                                            # :103
    ef4096 = roi_fmap_edge(ef, p)                                                   # :104
    return l1_forward(nf4096, ef4096, rel_inds[:, 1:3], p, mp_iter)                 # :103-107


# --------------------------------------------------------------------------- #
# a11 / a15: host-side helpers
# --------------------------------------------------------------------------- #
def get_rel_inds_eval(im_inds):
    """RelModelBase.get_rel_inds, eval branch without overlap filter
    (rel_model_base.py:148-163): all same-image ordered pairs i != j as (img,i,j),
    in row-major (i, j) order (torch.nonzero order)."""
    im_inds = np.asarray(im_inds)
    cand = im_inds[:, None] == im_inds[None]
    np.fill_diagonal(cand, False)
    ij = np.argwhere(cand)
    return np.concatenate((im_inds[ij[:, 0]][:, None], ij), 1).astype(np.int64)


def softmax(x, axis=1):
    m = x.max(axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis, keepdims=True)


def filter_dets(boxes, obj_scores, obj_classes, rel_inds, pred_scores):
    """lib/surgery.py:17-55: triple score = max_{p>=1} pred[p] * s_subj * s_obj,
    sort descending.  (torch.sort is unstable; ties are ordered by a stable
    argsort here, tests compare tie-free or as sets within ties.)"""
    s0 = obj_scores[rel_inds[:, 0]]; s1 = obj_scores[rel_inds[:, 1]]
    score = pred_scores[:, 1:].max(1) * s0 * s1
    order = np.argsort(-score, kind='stable')
    return boxes, obj_classes, obj_scores, rel_inds[order], pred_scores[order]


# --------------------------------------------------------------------------- #
# training tail: losses, gradient clipping, SGD
# --------------------------------------------------------------------------- #
def cross_entropy_rows(logits, labels, ignore_index=-100):
    """torch.nn.functional.cross_entropy(reduction='none') restated: -log_softmax(x)[label];
    rows whose label equals ``ignore_index`` give 0."""
    x = logits.astype(np.float64)
    m = x.max(1, keepdims=True)
    lse = np.log(np.exp(x - m).sum(1)) + m[:, 0]
    keep = labels != ignore_index
    out = np.zeros(len(labels), np.float64)
    out[keep] = lse[keep] - x[np.nonzero(keep)[0], labels[keep]]
    return out


def edge_weights(rel_labels, loss_type, loss_weights=(1, 1, 1), idx_fg=None, idx_bg=None):
    """Row weights of lib/losses.py:36-62 (gamma folded in)."""
    alpha, beta, gamma = loss_weights
    if idx_fg is None:
        idx_fg = np.nonzero(rel_labels > 0)[0]                     # :27-28
    if idx_bg is None:
        idx_bg = np.nonzero(rel_labels == 0)[0]                    # :30-31
    M_FG, M_BG, M = len(idx_fg), len(idx_bg), len(rel_labels)      # :33
    if loss_type == 'baseline':
        assert alpha == beta == 1                                  # :42
        return np.full(M, float(gamma) / M)                        # :43
    if loss_type not in ('dnorm', 'dnorm-fgbg'):
        raise NotImplementedError(loss_type)                       # :66
    w = np.ones(M)                                                 # :48
    if M_FG > 0:
        w[idx_fg] = float(alpha) / M_FG                            # :51-52
    if loss_type == 'dnorm':
        if M_BG > 0 and M_FG > 0:
            w[idx_bg] = float(beta) / M_FG                         # :55-58
    elif M_BG > 0:
        w[idx_bg] = float(beta) / M_BG                             # :59-61
    return gamma * w                                               # :63


def edge_losses(rel_dists, rel_labels, loss_type='dnorm', loss_weights=(1, 1, 1), idx_fg=None, idx_bg=None,
                return_grad=False):
    """lib/losses.py:5-70 -> rel_loss (float64 scalar) [, d rel_loss / d rel_dists]."""
    w = edge_weights(rel_labels, loss_type, loss_weights, idx_fg, idx_bg)
    loss = float((w * cross_entropy_rows(rel_dists, rel_labels)).sum())
    if not return_grad:
        return loss
    g = softmax(rel_dists.astype(np.float64), 1)
    g[np.arange(len(rel_labels)), rel_labels] -= 1.0
    return loss, g * w[:, None]


def node_losses(rm_obj_dists, rm_obj_labels, return_grad=False):
    """lib/losses.py:73-74: CE with the default 'mean' reduction (mean over rows whose label is not -100)."""
    keep = rm_obj_labels != -100
    n = max(int(keep.sum()), 1)
    loss = float(cross_entropy_rows(rm_obj_dists, rm_obj_labels).sum() / n)
    if not return_grad:
        return loss
    g = softmax(rm_obj_dists.astype(np.float64), 1)
    g[np.nonzero(keep)[0], rm_obj_labels[keep]] -= 1.0
    g[~keep] = 0.0
    return loss, g / n


def clip_grad_norm(grads, max_norm, clip=True):
    """lib/pytorch_misc.py:625-664 on a list of gradient arrays (modified in place when clipped).
    -> (total_norm, clip_coef)."""
    total = 0.0
    for g in grads:
        total += float(np.sqrt((g.astype(np.float64) ** 2).sum())) ** 2        # :643-644
    total = total ** 0.5                                                         # :648
    coef = float(max_norm) / (total + 1e-6)                                      # :649
    if coef < 1 and clip:
        for g in grads:
            g *= np.float32(coef)                                                # :650-653
    return total, coef


def sgd_step(params, grads, bufs, lrs, weight_decay, momentum=0.9):
    """torch.optim.SGD (dampening 0, no Nesterov), the optimizer lib/pytorch_misc.py:146 builds:
    d = g + wd * p;  buf = d on the first step else momentum * buf + d;  p -= lr * buf.  float32, in place;
    ``bufs[i] is None`` marks a first step.  Returns the new buffer list."""
    out = []
    for p, g, b, lr in zip(params, grads, bufs, lrs):
        if g is None:
            out.append(b)
            continue
        d = g.astype(np.float32)
        if weight_decay != 0:
            d = d + np.float32(weight_decay) * p
        b = d.copy() if b is None else np.float32(momentum) * b + d
        p -= np.float32(lr) * b
        out.append(b)
    return out
