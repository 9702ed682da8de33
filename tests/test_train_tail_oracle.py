"""CPU: pins the numpy restatement of the training tail (lib/losses.py, clip_grad_norm, the SGD of get_optim) against
tests/golden/train_tail.npz, produced by running the reference (tests/golden/make_golden_train.py)."""
import numpy as np
import pytest
from oracle import imp_numpy as O
from sgg_b200 import synth
from tests import cases

FX = cases.load('train_tail')
RTOL = 2e-6     # float64 oracle vs the reference's float32 torch arithmetic


def case_inputs(name):
    c = synth.loss_cases()[name]
    logits, labels = synth.synth_logits(c['M'], c['C'], c['seed'], fg=c['fg'])
    assert synth.digest(logits, labels) == str(FX['digest_' + name]), 'generator drift'
    kw = {}
    if c.get('explicit_idx'):
        kw['idx_fg'], kw['idx_bg'] = synth.explicit_idx(labels, c['seed'])
    return c, logits, labels, kw


@pytest.mark.parametrize('name', sorted(synth.loss_cases()))
def test_losses_vs_reference(name):
    c, logits, labels, kw = case_inputs(name)
    if c['kind'] == 'node':
        loss, g = O.node_losses(logits, labels, return_grad=True)
    else:
        loss, g = O.edge_losses(logits, labels, c['kind'], c['w'], return_grad=True, **kw)
    ref = float(FX['loss_' + name])
    assert abs(loss - ref) <= RTOL * max(1.0, abs(ref)), (loss, ref)
    rows = FX['dlogits_rows_' + name]
    gscale = max(np.abs(FX['dlogits_' + name]).max(), 1e-12)
    assert np.abs(g[rows] - FX['dlogits_' + name]).max() <= 1e-5 * gscale
    assert np.abs(g.sum(0) - FX['dlogits_colsum_' + name]).max() <= 1e-5 * gscale * max(1, len(labels)) ** 0.5


def test_edge_loss_rejects_what_the_reference_rejects():
    logits, labels = synth.synth_logits(8, 51, 1)
    with pytest.raises(AssertionError):
        O.edge_losses(logits, labels, 'baseline', (2, 1, 1))
    with pytest.raises(NotImplementedError):
        O.edge_losses(logits, labels, 'focal')


def test_clip_and_sgd_vs_reference():
    tt = synth.synth_train_tail(seed=5)
    names = list(tt['params'])
    assert synth.digest(*[tt['params'][k] for k in names],
                        *[g for s in tt['grads'] for g in s.values() if g is not None]) == str(FX['digest_train_tail'])
    params = [tt['params'][k].copy() for k in names]
    lrs = [tt['lr'] / 10.0 if k.startswith('roi_fmap') else tt['lr'] for k in names]     # lib/pytorch_misc.py:135-142
    assert sorted(set(lrs)) == sorted(FX['group_lrs'].tolist())
    bufs = [None] * len(names)
    for step in range(tt['steps']):
        grads = [None if tt['grads'][step][k] is None else tt['grads'][step][k].copy() for k in names]
        total, coef = O.clip_grad_norm([g for g in grads if g is not None], tt['clip'], clip=True)
        assert abs(total - FX['norms'][step]) <= 1e-6 * FX['norms'][step]
        if step == 0:
            assert coef < 1
            for k, g in zip(names, grads):
                np.testing.assert_allclose(g, FX['clipped_grad0_' + k], rtol=2e-6, atol=1e-9)
        bufs = O.sgd_step(params, grads, bufs, lrs, tt['l2'], 0.9)
        if step in (0, tt['steps'] - 1):
            for k, p in zip(names, params):
                np.testing.assert_allclose(p, FX['p_step%d_%s' % (step, k)], rtol=2e-6, atol=2e-8, err_msg=k)
    for k, b in zip(names, bufs):
        np.testing.assert_allclose(b, FX['m_final_' + k], rtol=2e-6, atol=2e-8, err_msg=k)
