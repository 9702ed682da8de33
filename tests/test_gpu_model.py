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
