"""GPU: model-level parity (L2 predict, L3 forward) against golden outputs of the reference's own
predict / forward, plus the training-mode contract (Result fields, gradients for all trainable tensors)."""
import numpy as np
import pytest
import torch

from tests import cases
from tests.golden.make_golden import l3_case
from sgg_b200 import synth

pytestmark = pytest.mark.gpu


class FakeData:
    ind_to_classes = ['__background__'] + ['c%d' % i for i in range(150)]
    ind_to_predicates = ['__background__'] + ['p%d' % i for i in range(50)]


@pytest.fixture(scope='module')
def model():
    from sgg_b200.model import RelModelStanford
    with torch.device('cuda'):        # random init on the device: the weights are overwritten by the fixtures anyway
        m = RelModelStanford(train_data=FakeData(), mode='predcls')
    m = m.cuda()
    m.eval()
    return m


def load(m, p):
    sd = m.state_dict()
    for k, v in p.items():
        sd[k].copy_(torch.from_numpy(v))


def test_l2_predict_vs_golden(model):
    fx = cases.load('l2_predict')
    nfe, efe, rel_inds, rois, p = cases.l2_inputs(fx)
    load(model, p)
    with torch.no_grad():
        od, rd = model.predict(torch.from_numpy(nfe).cuda(), torch.from_numpy(efe).cuda(),
                               torch.from_numpy(rel_inds).cuda(), torch.from_numpy(rois).cuda(), None)
    assert np.abs(od.cpu().numpy() - fx['obj_dists']).max() <= 1e-4
    assert np.abs(rd.cpu().numpy() - fx['rel_dists']).max() <= 1e-4


def test_l2_train_mode_predict_and_all_gradients_vs_reference_autograd(model):
    """predict() in TRAINING mode (batch-statistics BN in the geometry branch, fc6 on pools + broadcast geometry through the
    7x7-summed-weight backward, fc7, unary, 3 x fused message passing with the tape, heads) and the gradients of all 40
    trainable tensors against the reference's own forward + torch.autograd (tests/golden/l2_train_grad.npz; dropout
    probability 0 on both sides)."""
    fx = cases.load('l2_train_grad')
    seed = int(fx['seed'])
    g = synth.synth_graph(3, 6, 14, seed)
    N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
    nfe, efe = synth.synth_pooled(N, E, seed)
    p = synth.synth_params(seed, level='l2', scale=1.0)
    assert synth.digest(nfe, efe, p['roi_fmap.1.0.weight'][:64]) == str(fx['digest']), 'generator drift'
    load(model, p)
    for bn in (model.union_boxes.conv[2], model.union_boxes.conv[6]):
        bn.num_batches_tracked.zero_()
    model.train()
    old_p, model.dropout_p = model.dropout_p, 0.0
    try:
        for q in model.parameters():
            q.grad = None
        rng = np.random.default_rng(seed + 5)
        r1 = torch.from_numpy(rng.standard_normal((N, 151), dtype=np.float32)).cuda()
        r2 = torch.from_numpy(rng.standard_normal((E, 51), dtype=np.float32)).cuda()
        od, rd = model.predict(torch.from_numpy(nfe).cuda(), torch.from_numpy(efe).cuda(), torch.from_numpy(g['rel_inds']).cuda(),
                               torch.from_numpy(g['rois']).cuda(), None)
        loss = (od * r1).sum() + (rd * r2).sum()
        loss.backward()
    finally:
        model.dropout_p = old_p
        model.eval()
    assert np.abs(od.detach().cpu().numpy() - fx['obj_dists']).max() <= 1e-4
    assert np.abs(rd.detach().cpu().numpy() - fx['rel_dists']).max() <= 1e-4
    assert abs(float(loss) - float(fx['loss'])) <= 2e-3
    assert np.abs(model.union_boxes.conv[2].running_mean.cpu().numpy() - fx['rm1']).max() <= 1e-6
    assert np.abs(model.union_boxes.conv[6].running_var.cpu().numpy() - fx['rv2']).max() <= 1e-6
    params = dict(model.named_parameters())
    names = [str(n) for n in fx['names']]
    assert len(names) == 40
    for k in names:
        gq = params[k].grad
        assert gq is not None, 'no gradient for ' + k
        kk = k.replace('.', '__')
        flat = gq.detach().cpu().numpy().reshape(-1)
        ref = fx['val__' + kk]; got = flat[fx['idx__' + kk]]
        asum_ref = float(fx['asum__' + kk])
        # scale: the tensor's mean |gradient| (the sampled entries of the huge fc6 matrices can all be tiny)
        scale = max(float(np.abs(ref).max()), asum_ref / flat.shape[0], 1e-12)
        err = float(np.abs(got - ref).max()) / scale
        assert err <= 5e-4, '%s: max|d|/scale = %.3e' % (k, err)
        asum = float(np.abs(flat).astype(np.float64).sum())
        assert abs(asum - asum_ref) <= 5e-4 * max(asum_ref, 1e-12), k


@pytest.mark.parametrize('mode', ['predcls', 'sgcls'])
def test_l3_forward_eval_vs_golden(model, mode):
    fx = cases.load('l3_forward')
    seed = int(fx['seed'])
    sizes, boxes, gt_classes, gt_rels = l3_case(seed)
    imgs = synth.synth_images(sizes, seed)
    p = synth.synth_params(seed, level='l2', scale=1.0)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith('detector.backbone.')}
    p.update(synth.synth_backbone_params(shapes, seed))
    assert synth.digest(imgs[0], imgs[1], boxes, p['detector.backbone.0.weight']) == str(fx['digest'])
    load(model, p)
    model.mode = mode
    model.eval()
    batch = [([torch.from_numpy(i)[None] for i in imgs], None, 0, torch.from_numpy(boxes), torch.from_numpy(gt_classes),
              torch.from_numpy(gt_rels), None, ['a', 'b'])]
    with torch.no_grad():
        b, oc, osc, rels, ps = model(batch)
    model.mode = 'predcls'
    assert np.abs(b - fx[mode + '_boxes']).max() <= 1e-3
    assert np.array_equal(oc, fx[mode + '_obj_classes'])
    assert np.abs(osc - fx[mode + '_obj_scores']).max() <= 1e-4
    # ranking: identical order except where the reference's own scores are closer than the tolerance
    ref_rels, ref_ps = fx[mode + '_rels'], fx[mode + '_pred_scores']
    assert rels.shape == ref_rels.shape and ps.shape == ref_ps.shape
    key = lambda r: {tuple(x): i for i, x in enumerate(r.tolist())}
    mine, theirs = key(rels), key(ref_rels)
    assert set(mine) == set(theirs)
    for pair, i in theirs.items():
        assert np.abs(ps[mine[pair]] - ref_ps[i]).max() <= 1e-4, pair
    same = (rels == ref_rels).all(1).mean()
    assert same >= 0.9, same


def test_l3_forward_eval_without_no_grad_matches_the_no_grad_path(model):
    """model.eval(); model(batch) with autograd enabled (a caller who forgot torch.no_grad()) takes the fp32 activation
    path — the pre-split fc6 / fc7 kernels cannot enter an autograd graph — and returns the same detections."""
    seed = 4242
    sizes, boxes, gt_classes, gt_rels = l3_case(seed)
    imgs = synth.synth_images(sizes, seed)
    model.mode = 'sgcls'
    model.eval()
    batch = [([torch.from_numpy(i)[None] for i in imgs], None, 0, torch.from_numpy(boxes), torch.from_numpy(gt_classes),
              torch.from_numpy(gt_rels), None, ['a', 'b'])]
    with torch.no_grad():
        b0, oc0, osc0, rels0, ps0 = model(batch)
    b1, oc1, osc1, rels1, ps1 = model(batch)
    model.mode = 'predcls'
    assert np.array_equal(b0, b1) and np.array_equal(oc0, oc1) and np.abs(osc0 - osc1).max() <= 1e-5
    assert rels0.shape == rels1.shape and ps0.shape == ps1.shape
    key = lambda r: {tuple(x): i for i, x in enumerate(r.tolist())}
    m0, m1 = key(rels0), key(rels1)
    assert set(m0) == set(m1)
    for pair, i in m0.items():
        assert np.abs(ps0[i] - ps1[m1[pair]]).max() <= 1e-4, pair


def test_train_forward_backward_contract(model):
    """Training mode returns a Result with the reference's fields; loss.backward() yields gradients for every
    trainable (non-detector) tensor; detector stays frozen."""
    seed = 91
    torch.manual_seed(0); np.random.seed(0)
    sizes, boxes, gt_classes, gt_rels = l3_case(8235)
    imgs = synth.synth_images(sizes, seed)
    for n, prm in model.detector.named_parameters():
        prm.requires_grad = False                       # main.py:62-63
    model.train()
    batch = [([torch.from_numpy(i)[None] for i in imgs], None, 0, torch.from_numpy(boxes).cuda(),
              torch.from_numpy(gt_classes).cuda(), torch.from_numpy(gt_rels).cuda(), None, None, ['a', 'b'])]
    res = model(batch)
    for f in ('rm_obj_dists', 'rel_dists', 'rm_obj_labels', 'rel_labels', 'rel_inds', 'im_inds', 'rois', 'rm_box_priors',
              'rm_box_priors_org', 'node_feat', 'edge_feat', 'fmap', 'im_sizes', 'im_sizes_org'):
        assert hasattr(res, f), f
    N, E = res.rm_obj_dists.shape[0], res.rel_dists.shape[0]
    assert res.rm_obj_dists.shape == (N, 151) and res.rel_dists.shape == (E, 51) and res.rel_labels.shape == (E, 4)
    assert not res.fmap.requires_grad
    loss = torch.nn.functional.cross_entropy(res.rm_obj_dists, res.rm_obj_labels) + \
        torch.nn.functional.cross_entropy(res.rel_dists, res.rel_labels[:, 3])
    model.zero_grad()
    loss.backward()
    missing = [n for n, prm in model.named_parameters() if not n.startswith('detector.') and prm.grad is None]
    assert not missing, missing
    assert all(prm.grad is None for n, prm in model.detector.named_parameters())
    assert all(torch.isfinite(prm.grad).all() for n, prm in model.named_parameters() if prm.grad is not None)
    model.eval()
