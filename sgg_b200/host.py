"""Host-side logic of the relation-model path (index bookkeeping, edge sampling, ranking).

Semantics follow the reference functions cited in each docstring; implementations are
vectorised torch/numpy written for this package (no per-image Python loops where the
reference has them, no forced host syncs except where the reference's contract returns
numpy).  None of this is arithmetic-heavy; the heavy stages live in ``sgg_b200.ops``.
"""
import numpy as np
import torch

REL_FG_FRACTION = 0.25      # config.py:34
IM_SCALE = 592              # config.py:32
BATCHNORM_MOMENTUM = 0.01   # config.py:35


class Result(object):
    """Attribute bag returned by ``forward`` in training mode; ``None`` fields are dropped
    (lib/pytorch_misc.py:682-700 does this "for WandB")."""

    _FIELDS = ('od_obj_dists', 'rm_obj_dists', 'obj_scores', 'obj_preds', 'obj_fmap', 'od_box_deltas',
               'rm_box_deltas', 'od_box_targets', 'rm_box_targets', 'od_box_priors', 'rm_box_priors',
               'boxes_assigned', 'boxes_all', 'od_obj_labels', 'rm_obj_labels', 'rpn_scores', 'rpn_box_deltas',
               'rel_labels', 'rel_labels_all', 'im_inds', 'fmap', 'rel_dists', 'rel_inds', 'rel_rep')

    def __init__(self, **kw):
        for k, v in kw.items():
            if k not in self._FIELDS:
                raise TypeError('unexpected Result field %r' % k)
            if v is not None:
                setattr(self, k, v)

    def is_none(self):
        return len(self.__dict__) == 0

    def __getitem__(self, index):
        d = self.__dict__
        return [d[k] for k in sorted(d.keys())][index]


def image_segments(im_inds):
    """[(img, start, end)] for a sorted image-index vector (lib/pytorch_misc.py:493-502
    ``enumerate_by_image``).  One host sync (the reference has one per call too)."""
    if im_inds.numel() == 0:
        return []
    vals, counts = torch.unique_consecutive(im_inds.long(), return_counts=True)
    vals = vals.tolist(); ends = torch.cumsum(counts, 0).tolist()
    out, s = [], 0
    for v, e in zip(vals, ends):
        out.append((int(v), s, int(e)))
        s = int(e)
    return out


def enumerate_by_image(im_inds):
    for seg in image_segments(im_inds):
        yield seg


def box_iou(a, b):
    """torchvision.ops.box_iou restated (used through lib/pytorch_misc.py:60-67 bbox_overlaps)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None] - inter)


def get_rel_inds(im_inds, rel_labels=None, training=False, box_priors=None, require_overlap=False):
    """RelModelBase.get_rel_inds (sgg_models/rel_model_base.py:143-165).
    train: rel_labels[:, :3]; eval: all same-image ordered pairs i != j (row-major order), optionally
    only overlapping boxes (SGDet).  Returns int64 [E,3] = (img, subj, obj) with GLOBAL object ids."""
    if training:
        return rel_labels[:, :3].detach().clone()
    same = im_inds[:, None] == im_inds[None]
    same.fill_diagonal_(False)
    if require_overlap:
        same = same & (box_iou(box_priors.float(), box_priors.float()) > 0)
    pairs = same.nonzero()
    # no candidate pair (single-box images): the reference's `rel_cands.dim() == 0` dummy-pair fallback (:160-161) is
    # torch-0.3 legacy — nonzero() of an all-false mask is [0, 2] on every torch the reference supports — so it returns
    # an empty [0, 3]; so do we (every kernel accepts E = 0)
    return torch.cat((im_inds[pairs[:, 0]][:, None], pairs), 1)


def union_rois(rois, union_inds):
    """sgg_models/rel_model_base.py:248-250 (kept for API users; the CUDA RoIAlign computes it on the fly)."""
    a, b = rois[union_inds[:, 0]], rois[union_inds[:, 1]]
    return torch.cat((a[:, :1], torch.min(a[:, 1:3], b[:, 1:3]), torch.max(a[:, 3:5], b[:, 3:5])), 1)


def filter_dets(boxes, obj_scores, obj_classes, rel_inds, pred_scores, logits=False):
    """lib/surgery.py:17-55.  Ranks candidate edges by max_{p>=1} P(p) * s_subj * s_obj (descending) and
    returns the reference's 5 numpy arrays.  CUDA tensors: scoring (+ the softmax when ``logits``), sort and gather
    run in libsgg_b200.so (csrc/eval_tail.cu, ties broken by edge id).  There is no CPU path: CPU tensors raise
    ``SggError`` (argument validation happens first, as in the reference)."""
    if boxes.dim() != 2:
        raise ValueError('Boxes needs to be [num_box, 4] but its {}'.format(tuple(boxes.shape)))
    assert obj_scores.shape[0] == boxes.shape[0], (obj_scores.shape, boxes.shape)
    assert rel_inds.shape[1] == 2 and pred_scores.shape[0] == rel_inds.shape[0]
    from . import ops
    rels, ps, _, _ = ops.rank_relations(pred_scores, obj_scores.float(), rel_inds, logits=logits)   # raises on CPU tensors
    return (boxes.detach().cpu().numpy(), obj_classes.detach().cpu().numpy(), obj_scores.detach().cpu().numpy(),
            rels.cpu().numpy(), ps.cpu().numpy())


def random_choose(t, num, p=None):
    """lib/pytorch_misc.py:555-570 (numpy RNG, as the reference) — device-agnostic here."""
    if min(t.shape[0], num) == t.shape[0]:
        return t
    idx = np.random.choice(t.shape[0], size=num, replace=False, p=p)
    return t[torch.as_tensor(idx, dtype=torch.long, device=t.device)].contiguous()


def proposal_assignments_gtbox(rois, gt_boxes, gt_classes, gt_rels, image_offset, RELS_PER_IMG, sample_factor=-1):
    """Edge sampler for PredCls/SGCls training (lib/proposal_assignments_gtbox.py:6-80).

    FG = annotated relations (capped at RELS_PER_IMG*0.25*num_im), BG = all other same-image ordered
    pairs, filling up to RELS_PER_IMG*num_im (or num_fg*sample_factor); rows (img, subj, obj, pred)
    with GLOBAL object ids, sorted by (img, subj, obj)."""
    im_inds = rois[:, 0].long()
    n = im_inds.shape[0]
    num_im = int(im_inds[-1].item()) + 1
    # local -> global object ids: offset of the first box of every image
    first = torch.zeros(num_im, dtype=torch.long, device=im_inds.device)
    first.scatter_reduce_(0, im_inds, torch.arange(n, device=im_inds.device), reduce='amin', include_self=False)
    fg = gt_rels.clone()
    fg[:, 0] -= image_offset
    fg[:, 1:3] += first[fg[:, 0]][:, None]

    cand = im_inds[:, None] == im_inds[None]
    cand.fill_diagonal_(False)
    cand[fg[:, 1], fg[:, 2]] = False
    bg_pairs = cand.nonzero()

    num_fg = min(fg.shape[0], int(RELS_PER_IMG * REL_FG_FRACTION * num_im))
    if num_fg < fg.shape[0]:
        fg = random_choose(fg, num_fg)
    sample_bg = (num_im > 1) and sample_factor > -1
    num_bg = min(bg_pairs.shape[0], int(num_fg * sample_factor) if sample_bg else int(RELS_PER_IMG * num_im) - num_fg)
    if num_bg > 0:
        bg = torch.cat((im_inds[bg_pairs[:, 0]][:, None], bg_pairs, torch.zeros_like(bg_pairs[:, :1])), 1)
        if num_bg < bg_pairs.shape[0]:
            bg = random_choose(bg, num_bg)
        rel_labels = torch.cat((fg, bg), 0)
    else:
        rel_labels = fg
    key = rel_labels[:, 0] * (n ** 2) + rel_labels[:, 1] * n + rel_labels[:, 2]
    rel_labels = rel_labels[torch.sort(key)[1]].contiguous()
    return rois, gt_classes[:, 1].contiguous(), rel_labels


def rel_assignments(im_inds, rpn_rois, roi_gtlabels, gt_boxes, gt_classes, gt_rels, image_offset,
                    fg_thresh=0.5, num_sample_per_gt=4, filter_non_overlap=True):
    """Edge sampler for SGDet training (lib/rel_assignments.py:12-137): detections are matched to GT boxes by
    class and IoU >= fg_thresh; FG relations are sampled per GT relation proportionally to the IoU product
    (<= 16 per image), BG pairs (overlapping, non-identical, both labelled) fill up to 64 per image.
    Returns int64 [R,4] = (img, subj, obj, pred) on the device of ``rpn_rois``."""
    fg_per_im = int(np.round(REL_FG_FRACTION * 64))
    dev = rpn_rois.device
    pi = im_inds.cpu().numpy(); pb = rpn_rois.detach().cpu().numpy(); pl = roi_gtlabels.cpu().numpy()
    gb = gt_boxes.detach().cpu().numpy(); gc = gt_classes.cpu().numpy().copy(); gr = gt_rels.cpu().numpy().copy()
    gc[:, 0] -= image_offset; gr[:, 0] -= image_offset
    num_im = int(gc[:, 0].max()) + 1
    iou = lambda a, b: box_iou(torch.from_numpy(a).float(), torch.from_numpy(b).float()).numpy()
    out, seen = [], 0
    for im in range(num_im):
        sel = np.where(pi == im)[0]
        gsel = np.where(gc[:, 0] == im)[0]
        boxes_i, labels_i = pb[sel], pl[sel]
        rels_i = gr[gr[:, 0] == im, 1:]
        ious = iou(boxes_i, gb[gsel])
        match = (labels_i[:, None] == gc[gsel, 1][None]) & (ious >= fg_thresh)
        self_iou = iou(boxes_i, boxes_i)
        overlap = (self_iou < 1) & (self_iou > 0)
        nb = boxes_i.shape[0]
        poss = overlap.copy() if filter_non_overlap else (np.ones((nb, nb), np.int64) - np.eye(nb, dtype=np.int64))
        poss[labels_i == 0] = 0
        poss[:, labels_i == 0] = 0
        fg = []
        for s_gt, o_gt, pred in rels_i:
            cands, scores = [], []
            for a in np.where(match[:, s_gt])[0]:
                for b in np.where(match[:, o_gt])[0]:
                    if a != b:
                        cands.append((a, b, pred)); scores.append(ious[a, s_gt] * ious[b, o_gt])
                        poss[a, b] = 0
            if not cands:
                continue
            pr = np.asarray(scores); pr = pr / pr.sum()
            for j in np.random.choice(len(cands), p=pr, size=min(len(cands), num_sample_per_gt), replace=False):
                fg.append(cands[j])
        fg = np.asarray(fg, dtype=np.int64).reshape(-1, 3)
        if fg.shape[0] > fg_per_im:
            fg = fg[np.random.choice(fg.shape[0], size=fg_per_im, replace=False)]
        bg = np.column_stack(np.where(poss))
        bg = np.column_stack((bg, np.zeros(bg.shape[0], dtype=np.int64)))
        if bg.size > 0:
            bg = bg[np.random.choice(bg.shape[0], size=min(64 - fg.shape[0], bg.shape[0]), replace=False)]
        else:
            bg = np.zeros((0, 3), dtype=np.int64)
        if fg.size == 0 and bg.size == 0:
            bg = np.array([[0, 0, 0]], dtype=np.int64)
        rows = np.concatenate((fg, bg), 0)
        rows[:, 0:2] += seen
        rows = rows[np.lexsort((rows[:, 1], rows[:, 0]))]
        out.append(np.column_stack((np.full(rows.shape[0], im, np.int64), rows)))
        seen += nb
    return torch.from_numpy(np.concatenate(out, 0)).to(dev)


def dataset_counts(train_data, must_overlap=True):
    """Relation statistics for the frequency baseline (lib/get_dataset_counts.py:10-66): fg[o1,o2,pred] counts the
    annotated triples; bg[o1,o2] counts ordered pairs of (overlapping, if ``must_overlap``; all pairs when an image has
    no overlapping boxes) GT boxes.  ``train_data`` exposes num_classes, num_predicates and per-image lists
    gt_classes / relationships / gt_boxes like the reference's VG dataset.  Vectorised with np.add.at."""
    C, R = train_data.num_classes, train_data.num_predicates
    fg = np.zeros((C, C, R), dtype=np.int64)
    bg = np.zeros((C, C), dtype=np.int64)
    for i in range(len(train_data)):
        cls = np.asarray(train_data.gt_classes[i])
        rels = np.asarray(train_data.relationships[i])
        boxes = torch.as_tensor(np.asarray(train_data.gt_boxes[i]), dtype=torch.float64)
        if rels.size:
            np.add.at(fg, (cls[rels[:, 0]], cls[rels[:, 1]], rels[:, 2]), 1)
        n = cls.shape[0]
        pairs = ~np.eye(n, dtype=bool)
        if must_overlap:
            ov = (box_iou(boxes, boxes) > 0).numpy() & pairs
            if ov.any():
                pairs = ov
        a, b = np.nonzero(pairs)
        np.add.at(bg, (cls[a], cls[b]), 1)
    return fg, bg


class FrequencyBias(torch.nn.Module):
    """log P(pred | subj, obj) lookup added to rel_dists (lib/sparse_targets.py:7-33).
    Built from (fg_matrix [C,C,R], bg_matrix [C,C]) count arrays — the caller computes them from its
    dataset (lib/get_dataset_counts.py); state-dict key ``obj_baseline.weight`` as in the reference."""

    def __init__(self, fg_matrix, bg_matrix, eps=1e-3):
        super().__init__()
        fg = np.array(fg_matrix, dtype=np.float64).copy()
        fg[:, :, 0] = np.asarray(bg_matrix, dtype=np.float64) + 1
        dist = np.log(fg / fg.sum(2)[:, :, None] + eps)
        self.num_objs = dist.shape[0]
        w = torch.tensor(dist.reshape(-1, dist.shape[2]), dtype=torch.float32)
        self.obj_baseline = torch.nn.Embedding(w.shape[0], w.shape[1])
        self.obj_baseline.weight.data = w

    def index_with_labels(self, labels):
        return self.obj_baseline(labels[:, 0] * self.num_objs + labels[:, 1])

    def forward(self, obj_cands0, obj_cands1):
        """lib/sparse_targets.py:36-50: expected log-frequency under two class distributions [B,C] -> [B,R]."""
        joint = obj_cands0[:, :, None] * obj_cands1[:, None]
        return joint.view(joint.size(0), -1) @ self.obj_baseline.weight
