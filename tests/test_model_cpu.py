"""CPU: drop-in contract of the model class and host-side logic (no CUDA compute)."""
import json
import os
import numpy as np
import pytest
import torch

from tests import cases
from sgg_b200 import host


class FakeData:
    ind_to_classes = ['__background__'] + ['c%d' % i for i in range(150)]
    ind_to_predicates = ['__background__'] + ['p%d' % i for i in range(50)]


@pytest.fixture(scope='module')
def model():
    from sgg_b200.model import RelModelStanford
    return RelModelStanford(train_data=FakeData(), mode='predcls')


def test_state_dict_keys_and_shapes_match_reference(model):
    ref = json.load(open(os.path.join(cases.GOLDEN, 'state_dict_keys.json')))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert sorted(mine) == sorted(ref)
    for k in ref:
        assert mine[k] == ref[k], k
    # optimizer grouping of lib/pytorch_misc.py:135-142: names starting with roi_fmap go to the lr/10 group
    fc = [n for n, p in model.named_parameters() if n.startswith('roi_fmap')]
    assert len(fc) == 8
    assert sum(p.numel() for n, p in model.named_parameters() if not n.startswith('detector.')) == 247753678


def test_plugin_surface(model):
    for attr in ('detector', 'edge_dim', 'pool_sz', 'fmap_sz', 'mode', 'hidden_dim', 'mp_iter', 'roi_fmap',
                 'roi_fmap_obj', 'union_boxes', 'roi_pool'):
        assert hasattr(model, attr)
    assert (model.edge_dim, model.pool_sz, model.fmap_sz, model.hidden_dim, model.mp_iter) == (512, 7, 38, 512, 3)
    for fn in ('forward', 'predict', 'message_pass', 'node_edge_features', 'get_rel_inds', 'get_scaled_boxes',
               'set_box_score_thresh', 'faster_rcnn', 'gt_labels'):
        assert callable(getattr(model, fn))
    with pytest.raises(AssertionError):
        model([1, 2])                                   # len(batch) != 1 (rel_model_stanford.py:121)
    model.set_box_score_thresh(0.05)
    assert model.detector.roi_heads.score_thresh == 0.05


def test_get_rel_inds_eval_and_train():
    im = torch.tensor([0, 0, 0, 1, 1])
    r = host.get_rel_inds(im)
    from oracle import imp_numpy as O
    assert np.array_equal(r.numpy(), O.get_rel_inds_eval(im.numpy()))
    labels = torch.tensor([[0, 0, 1, 5], [0, 1, 2, 0]])
    assert host.get_rel_inds(im, labels, training=True).tolist() == [[0, 0, 1], [0, 1, 2]]
    boxes = torch.tensor([[0, 0, 10, 10], [5, 5, 20, 20], [50, 50, 60, 60], [0, 0, 5, 5], [1, 1, 3, 3]], dtype=torch.float32)
    r2 = host.get_rel_inds(im, box_priors=boxes, require_overlap=True)
    assert r2.tolist() == [[0, 0, 1], [0, 1, 0], [1, 3, 4], [1, 4, 3]]


def test_filter_dets_has_no_cpu_path():
    """Argument validation mirrors lib/surgery.py:17-55 (ValueError first); the ranking itself only exists as CUDA
    kernels, so CPU tensors are refused loudly instead of silently taking a torch path."""
    from sgg_b200._lib import SggError
    rng = np.random.default_rng(0)
    N, E = 6, 30
    boxes = rng.random((N, 4)).astype(np.float32); scores = rng.random(N).astype(np.float32)
    cls = rng.integers(1, 151, N); rel = np.stack((rng.integers(0, N, E), rng.integers(0, N, E)), 1)
    ps = rng.random((E, 51)).astype(np.float32)
    with pytest.raises(SggError):
        host.filter_dets(torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(cls),
                         torch.from_numpy(rel), torch.from_numpy(ps))
    with pytest.raises(ValueError):
        host.filter_dets(torch.zeros(2, 3, 4), torch.zeros(2), torch.zeros(2), torch.zeros(1, 2).long(), torch.zeros(1, 51))


def test_proposal_assignments_gtbox_semantics():
    from sgg_b200 import synth
    g = synth.synth_graph(3, 6, 10, 5)
    rois = torch.from_numpy(g['rois']); gt_cls = torch.from_numpy(g['gt_classes']); gt_rels = torch.from_numpy(g['gt_rels'])
    _, labels, rl = host.proposal_assignments_gtbox(rois, rois[:, 1:], gt_cls, gt_rels, 0, 1024)
    n = rois.shape[0]
    assert rl.shape[0] == 3 * 6 * 5                       # every ordered same-image pair exactly once
    key = rl[:, 0] * n * n + rl[:, 1] * n + rl[:, 2]
    assert torch.all(key[1:] > key[:-1])                  # sorted by (img, subj, obj), no duplicates
    assert torch.all(gt_cls[rl[:, 1], 0] == rl[:, 0]) and torch.all(gt_cls[rl[:, 2], 0] == rl[:, 0])
    assert int((rl[:, 3] > 0).sum()) == gt_rels.shape[0]
    fg = rl[rl[:, 3] > 0]
    first = torch.tensor([0, 6, 12])
    want = sorted((int(r[0]), int(r[1] + first[r[0]]), int(r[2] + first[r[0]]), int(r[3])) for r in gt_rels)
    assert sorted(map(tuple, fg.tolist())) == want
    # budget: RELS_PER_IMG * num_im total, FG capped at 25 %
    np.random.seed(0)
    _, _, rl2 = host.proposal_assignments_gtbox(rois, rois[:, 1:], gt_cls, gt_rels, 0, 8)
    assert rl2.shape[0] == 24 and int((rl2[:, 3] > 0).sum()) == 6


def test_result_container():
    r = host.Result(rel_dists=torch.zeros(2), od_obj_dists=None)
    assert hasattr(r, 'rel_dists') and not hasattr(r, 'od_obj_dists') and not r.is_none()
    with pytest.raises(TypeError):
        host.Result(bogus=1)


def test_rel_assignments_shapes():
    np.random.seed(1)
    boxes = torch.tensor([[0, 0, 50, 50], [10, 10, 60, 60], [40, 40, 90, 90], [0, 0, 30, 30], [20, 20, 80, 80]], dtype=torch.float32)
    im_inds = torch.tensor([0, 0, 0, 1, 1])
    labels = torch.tensor([3, 5, 7, 2, 9])
    gt_classes = torch.stack((im_inds, labels), 1)
    gt_rels = torch.tensor([[0, 0, 1, 4], [1, 0, 1, 6]])
    out = host.rel_assignments(im_inds, boxes, labels, boxes, gt_classes, gt_rels, 0, num_sample_per_gt=1)
    assert out.shape[1] == 4 and out.dtype == torch.int64
    assert [0, 0, 1, 4] in out.tolist() and [1, 3, 4, 6] in out.tolist()


def test_frequency_bias_counts_and_lookup():
    class DS:
        num_classes, num_predicates = 5, 4
        gt_classes = [np.array([1, 2, 3]), np.array([2, 2])]
        relationships = [np.array([[0, 1, 2], [1, 2, 3]]), np.array([[0, 1, 1]])]
        gt_boxes = [np.array([[0, 0, 10, 10], [5, 5, 20, 20], [50, 50, 60, 60]], float), np.array([[0, 0, 5, 5], [9, 9, 12, 12]], float)]
        def __len__(self): return 2
    fg, bg = host.dataset_counts(DS())
    assert fg[1, 2, 2] == 1 and fg[2, 3, 3] == 1 and fg[2, 2, 1] == 1 and fg.sum() == 3
    assert bg[1, 2] == 1 and bg[2, 1] == 1          # only the overlapping pair of image 0
    assert bg[2, 2] == 2                            # image 1 has no overlap -> all ordered pairs
    fb = host.FrequencyBias(fg, bg)
    assert fb.obj_baseline.weight.shape == (25, 4)
    out = fb.index_with_labels(torch.tensor([[1, 2], [2, 2]]))
    ref = np.log((np.array([bg[1, 2] + 1, 0, 1, 0]) / (bg[1, 2] + 1 + 1)) + 1e-3)
    assert np.allclose(out[0].detach().numpy(), ref, atol=1e-6)


@pytest.mark.skipif(not os.path.exists('/root/reference/lib/pytorch_misc.py'), reason='needs the reference checkout (build container)')
def test_reference_plumbing_drives_the_drop_in_model(model, tmp_path):
    """The reference's OWN helpers (lib/pytorch_misc.py: set_mode :76-95, get_optim :130-157, save_checkpoint :214-230,
    load_checkpoint -> optimistic_restore :160-211), imported from /root/reference and run unmodified on the sgg_b200
    model class: the drop-in claim of INTEGRATION.md section 1, executed."""
    import sys, types
    sys.path.insert(0, '/root/reference')
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))          # lib/pytorch_misc.py:7 imports it; not installed
    try:
        from lib import pytorch_misc as R
    finally:
        sys.path.remove('/root/reference')
    from sgg_b200.model import RelModelStanford
    from sgg_b200 import optim
    R.set_mode(model, 'sgcls', is_train=False)
    assert (model.mode, model.detector.mode, model.training) == ('sgcls', 'gtbox', False)
    R.set_mode(model, 'sgdet', is_train=True)
    assert (model.mode, model.detector.mode, model.training) == ('sgdet', 'refinerels', True)
    R.set_mode(model, 'predcls', is_train=False)
    conf = types.SimpleNamespace(l2=1e-4, steps=[15], lr_decay=0.1, ckpt='', device='cpu', gan=False, backbone='vgg16',
                                 mode='predcls')
    frozen = [p for p in model.detector.parameters() if p.requires_grad]
    for p in model.detector.parameters():                              # main.py:62-63
        p.requires_grad = False
    try:
        ref_opt, _ = R.get_optim(model, 0.01, conf, -1)                # the reference's grouping on our parameter names
        ours, _ = optim.get_optim(model, 0.01, conf, -1)
        assert [len(g['params']) for g in ref_opt.param_groups] == [len(g['params']) for g in ours.param_groups] == [8, 32]
        assert [g['lr'] for g in ref_opt.param_groups] == [g['lr'] for g in ours.param_groups]
        path = str(tmp_path / 'vgrel-3.tar')
        R.save_checkpoint(model, ours, path, {'epoch': 3, 'global_batch_iter': 77})
        other = RelModelStanford(train_data=FakeData(), mode='predcls')
        start_epoch, ckpt = R.load_checkpoint(conf, other, path)
        assert start_epoch == 3 and other.global_batch_iter == 77 and 'optimizer' in ckpt
        assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), other.state_dict().values()))
        os.remove(path)
    finally:
        for p in frozen:
            p.requires_grad = True
