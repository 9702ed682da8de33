"""GPU parity of the evaluation tail (csrc/eval_tail.cu through the C-ABI): ranking of candidate relations as
lib/surgery.py:17-55 filter_dets does, against the numpy oracle (O.filter_dets / O.softmax).  Index outputs are
compared exactly (inputs are tie-free unless a test says otherwise), probabilities to 1e-6."""
import numpy as np
import pytest
import torch

from oracle import imp_numpy as O

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make(N, E, seed, P=51):
    rng = np.random.default_rng(seed)
    logits = (2.0 * rng.standard_normal((E, P))).astype(np.float32)
    scores = rng.uniform(0.05, 1.0, N).astype(np.float32)
    rel = np.stack((rng.integers(0, N, E), rng.integers(0, N, E)), 1).astype(np.int64)
    return logits, scores, rel


@pytest.mark.parametrize('N,E', [(2, 1), (2, 2), (10, 90), (30, 870), (62, 3782), (64, 4096), (80, 5000), (300, 70000)])
def test_rank_relations_vs_oracle(N, E):
    from sgg_b200 import ops
    logits, scores, rel = make(N, E, 100 + E)
    rels, pred, sc, order = ops.rank_relations(dev(logits), dev(scores), dev(rel), logits=True, validate=True)
    probs = O.softmax(logits.astype(np.float64), 1).astype(np.float32)
    _, _, _, ref_rels, ref_pred = O.filter_dets(np.zeros((N, 4), np.float32), scores, np.zeros(N, np.int64), rel, probs)
    order = order.cpu().numpy()
    sc = sc.cpu().numpy()
    assert sorted(order.tolist()) == list(range(E))                       # a permutation
    assert np.all(sc[:-1] >= sc[1:])                                      # descending
    ref_score = probs[:, 1:].max(1) * scores[rel[:, 0]] * scores[rel[:, 1]]
    assert np.abs(sc - ref_score[order]).max() <= 1e-6
    # same ranking as the oracle wherever the oracle's scores are separated by more than fp32 noise
    ref_order = np.argsort(-ref_score, kind='stable')
    gap_ok = np.abs(np.diff(ref_score[ref_order])) > 1e-6
    same = order == ref_order
    bad = ~same
    if bad.any():       # any disagreement must sit inside a run of near-equal scores
        for i in np.nonzero(bad)[0]:
            lo, hi = max(i - 1, 0), min(i, E - 2)
            assert (not gap_ok[lo]) or (not gap_ok[hi]), 'rank %d differs outside a near-tie' % i
    assert np.array_equal(rels.cpu().numpy(), rel[order])
    assert np.abs(pred.cpu().numpy() - probs[order]).max() <= 1e-6
    if not bad.any():
        assert np.array_equal(rels.cpu().numpy(), ref_rels) and np.abs(pred.cpu().numpy() - ref_pred).max() <= 1e-6


def test_ties_are_broken_by_edge_id_and_probabilities_pass_through():
    from sgg_b200 import ops
    logits, scores, rel = make(12, 600, 7)
    probs = O.softmax(logits.astype(np.float64), 1).astype(np.float32)
    probs[100:400] = probs[5]                         # 301 identical rows ...
    rel[100:400] = rel[5]                             # ... with identical endpoints: exact score ties
    rels, pred, sc, order = ops.rank_relations(dev(probs), dev(scores), dev(rel), logits=False)
    order = order.cpu().numpy(); sc = sc.cpu().numpy()
    score = probs[:, 1:].max(1) * scores[rel[:, 0]] * scores[rel[:, 1]]
    assert np.array_equal(sc, score[order])           # no softmax, same multiplication order: bit-exact
    assert np.array_equal(order, np.argsort(-score, kind='stable'))
    assert np.array_equal(pred.cpu().numpy(), probs[order])


def test_per_image_ranking_equals_image_by_image():
    from sgg_b200 import ops
    rng = np.random.default_rng(3)
    sizes = [5, 1, 30, 17, 62]
    logits_all, rel_all, scores_all, base = [], [], [], 0
    for b, n in enumerate(sizes):
        E = max(n * (n - 1), 1)
        lg, sc, rl = make(n, E, 50 + b)
        logits_all.append(lg); scores_all.append(sc)
        rel_all.append(np.concatenate((np.full((E, 1), b, np.int64), rl + base), 1))
        base += n
    logits = np.concatenate(logits_all); scores = np.concatenate(scores_all); rel = np.concatenate(rel_all)
    perm = rng.permutation(len(rel))                  # rows need not arrive grouped by image
    rels, pred, sc, order = ops.rank_relations(dev(logits[perm]), dev(scores), dev(rel[perm]), logits=True, per_image=True,
                                               validate=True)
    rels = rels.cpu().numpy(); pred = pred.cpu().numpy()
    pos, base = 0, 0
    for b, n in enumerate(sizes):
        E = len(logits_all[b])
        r1, p1, _, _ = ops.rank_relations(dev(logits_all[b]), dev(scores), dev(rel_all[b][:, 1:]), logits=True)
        assert np.array_equal(rels[pos:pos + E], r1.cpu().numpy())
        assert np.array_equal(pred[pos:pos + E], p1.cpu().numpy())
        pos += E


def test_filter_dets_cuda_matches_oracle_and_rejects_bad_indices():
    from sgg_b200 import host, ops
    from sgg_b200._lib import SggError
    logits, scores, rel = make(20, 380, 11)
    boxes = np.random.default_rng(1).uniform(0, 500, (20, 4)).astype(np.float32)
    cls = np.arange(20, dtype=np.int64)
    from oracle import imp_numpy as O
    a = host.filter_dets(dev(boxes), dev(scores), dev(cls), dev(rel), dev(logits), logits=True)
    b = O.filter_dets(boxes, scores, cls, rel, O.softmax(logits.astype(np.float64), 1).astype(np.float32))
    for x, y in zip(a[:4], b[:4]):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    assert a[4].dtype == np.float32 and np.abs(a[4] - b[4]).max() <= 1e-6
    with pytest.raises(ValueError):
        host.filter_dets(dev(np.zeros((2, 3, 4), np.float32)), dev(scores[:2]), dev(cls[:2]), dev(rel[:1] * 0), dev(logits[:1]))
    bad = rel.copy(); bad[7, 1] = 20
    with pytest.raises(SggError):
        ops.rank_relations(dev(logits), dev(scores), dev(bad), validate=True)
    r, p, s, o = ops.rank_relations(dev(logits[:0]), dev(scores), dev(rel[:0]))
    assert r.shape == (0, 2) and p.shape == (0, 51)
