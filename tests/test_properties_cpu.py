"""CPU property tests (hypothesis) of the oracle and the host logic on ragged graphs — SURVEY.md §4 item 3: N = 2,
isolated nodes, duplicate edges, unsorted rel_inds.  These pin the invariants the GPU property tests then check on the
CUDA path at full size (tests/test_gpu_parity.py): edge-permutation equivariance, image independence, duplicates add."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle import imp_numpy as O
from sgg_b200 import host, synth

H = 64      # hidden size of the property runs (the algebra does not depend on it; 512 only costs time)
P = synth.synth_params(3, H=H, level='l0')
SET = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@st.composite
def graphs(draw):
    n_img = draw(st.integers(1, 3))
    sizes = [draw(st.integers(2, 6)) for _ in range(n_img)]
    rel, base = [], 0
    for b, n in enumerate(sizes):
        ne = draw(st.integers(0, 2 * n * (n - 1)))                      # may exceed n(n-1): duplicates on purpose
        for _ in range(ne):
            s = draw(st.integers(0, n - 1)); o = draw(st.integers(0, n - 1))
            if s != o:
                rel.append((b, base + s, base + o))
        base += n
    rel = np.array(rel, np.int64).reshape(-1, 3)
    seed = draw(st.integers(0, 2 ** 16))
    return sizes, rel, seed


def states(N, E, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((N, H)).astype(np.float32) * 0.5, np.maximum(rng.standard_normal((E, H)), 0).astype(np.float32))


@settings(**SET)
@given(graphs(), st.integers(1, 3))
def test_edge_permutation_equivariance_and_unsorted_edges(g, T):
    sizes, rel, seed = g
    N, E = sum(sizes), rel.shape[0]
    obj, er = states(N, E, seed)
    v, e = O.message_pass(er, obj, rel[:, 1:3], P, T)
    perm = np.random.default_rng(seed + 1).permutation(E)
    v2, e2 = O.message_pass(er[perm], obj, rel[perm][:, 1:3], P, T)
    assert np.abs(v - v2).max() <= 2e-5 if N else True
    assert E == 0 or np.abs(e[perm] - e2).max() <= 2e-5


@settings(**SET)
@given(graphs(), st.integers(1, 2))
def test_images_are_independent(g, T):
    """rel_inds only connect objects of one image (rel_model_base.py:148): a batch equals its images run one by one."""
    sizes, rel, seed = g
    N, E = sum(sizes), rel.shape[0]
    obj, er = states(N, E, seed)
    v, e = O.message_pass(er, obj, rel[:, 1:3], P, T)
    base = 0
    for b, n in enumerate(sizes):
        m = rel[:, 0] == b
        vb, eb = O.message_pass(er[m], obj[base:base + n], rel[m][:, 1:3] - base, P, T)
        assert np.abs(v[base:base + n] - vb).max() <= 2e-5
        assert m.sum() == 0 or np.abs(e[m] - eb).max() <= 2e-5
        base += n


@settings(**SET)
@given(st.integers(2, 6), st.integers(0, 2 ** 16))
def test_isolated_nodes_and_duplicate_edges(n, seed):
    """A node without edges is updated by node_gru with a zero message (rel_model_stanford.py:91-92); a duplicated edge
    contributes twice to its endpoints' messages and both copies get the same state."""
    rng = np.random.default_rng(seed)
    rel = np.array([[0, 1], [0, 1]], np.int64)                          # nodes >= 2 are isolated
    obj, er = states(n, 2, seed)
    er[1] = er[0]
    v, e = O.message_pass(er, obj, rel, P, 1)
    assert np.array_equal(e[0], e[1])
    v0 = O.gru_cell(obj, np.zeros_like(obj), P['node_gru.weight_ih'], P['node_gru.weight_hh'], P['node_gru.bias_ih'],
                    P['node_gru.bias_hh'])                              # V_0 = node_gru(obj_rep, h = 0)  (:72)
    if n > 2:
        iso = O.gru_cell(np.zeros((n - 2, H), np.float32), v0[2:], P['node_gru.weight_ih'], P['node_gru.weight_hh'],
                         P['node_gru.bias_ih'], P['node_gru.bias_hh'])
        assert np.abs(v[2:] - iso).max() <= 2e-5
    v1, _ = O.message_pass(er[:1], obj, rel[:1], P, 1)                  # without the duplicate the endpoints differ
    assert np.abs(v[:2] - v1[:2]).max() > 1e-7
    del rng


@settings(**SET)
@given(st.lists(st.integers(1, 7), min_size=1, max_size=4))
def test_get_rel_inds_eval_enumerates_same_image_ordered_pairs(sizes):
    im_inds = torch.from_numpy(np.repeat(np.arange(len(sizes)), sizes))
    rel = host.get_rel_inds(im_inds, None, training=False)
    assert rel.shape == (sum(n * (n - 1) for n in sizes), 3)
    if rel.shape[0]:
        assert torch.all(im_inds[rel[:, 1]] == rel[:, 0]) and torch.all(im_inds[rel[:, 2]] == rel[:, 0])
        assert torch.all(rel[:, 1] != rel[:, 2])
        key = rel[:, 1] * len(im_inds) + rel[:, 2]
        assert torch.all(key[1:] > key[:-1])                            # row-major (subject, object) order, no repeats
    assert np.array_equal(rel.numpy(), O.get_rel_inds_eval(im_inds.numpy()))


@settings(**SET)
@given(st.integers(1, 40), st.integers(2, 12), st.floats(0.0, 1.0), st.integers(0, 999))
def test_loss_weights_sum_rules(M, C, fg, seed):
    """lib/losses.py:36-62: 'baseline' weights sum to gamma; 'dnorm-fgbg' FG weights sum to alpha and BG weights to beta
    whenever the class is present; 'dnorm' divides both by M_FG."""
    rng = np.random.default_rng(seed)
    labels = (rng.random(M) < fg) * rng.integers(1, C, M)
    m_fg, m_bg = int((labels > 0).sum()), int((labels == 0).sum())
    assert abs(O.edge_weights(labels, 'baseline', (1, 1, 0.7)).sum() - 0.7) <= 1e-12
    w = O.edge_weights(labels, 'dnorm-fgbg', (2.0, 3.0, 1.0))
    if m_fg:
        assert abs(w[labels > 0].sum() - 2.0) <= 1e-9
    if m_bg:
        assert abs(w[labels == 0].sum() - 3.0) <= 1e-9
    w = O.edge_weights(labels, 'dnorm', (2.0, 3.0, 1.0))
    if m_fg:
        assert abs(w[labels > 0].sum() - 2.0) <= 1e-9 and (m_bg == 0 or abs(w[labels == 0].sum() - 3.0 * m_bg / m_fg) <= 1e-9)
    else:
        assert np.all(w == 1.0)
