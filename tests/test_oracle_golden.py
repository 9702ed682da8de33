"""Pins the numpy oracle against outputs of the reference itself (tests/golden,
produced by tests/golden/make_golden.py in the build container).  CPU only."""
import numpy as np
import pytest
from oracle import imp_numpy as O
from tests import cases

TOL = 2e-5   # fp32 reassociation noise between torch-CPU (MKL) and numpy; the product bar is 1e-4


@pytest.mark.parametrize('name', ['l0_cfg1', 'l0_cfg2_s3', 'l0_ragged_t6'])
def test_l0_message_pass(name):
    fx = cases.load(name)
    obj, rel, rel_inds, p, T = cases.l0_inputs(fx)
    v, e = O.message_pass(rel, obj, rel_inds[:, 1:3], p, T)
    cases.check_rows(v, fx['v_rows'], fx['v'], fx['v_colsum'], TOL, 'V')
    cases.check_rows(e, fx['e_rows'], fx['e'], fx['e_colsum'], TOL, 'E')


def test_l0_special_graph():
    fx = cases.load('l0_special')
    obj, rel, rel_inds, p, T = cases.l0_inputs(fx)
    v, e = O.message_pass(rel, obj, rel_inds[:, 1:3], p, T)
    assert np.abs(v - fx['v']).max() <= TOL and np.abs(e - fx['e']).max() <= TOL


@pytest.mark.parametrize('name', ['l1_cfg1', 'l1_cfg2', 'l1_cfg2_s3'])
def test_l1_forward(name):
    fx = cases.load(name)
    of, ef, rel_inds, p, T = cases.l1_inputs(fx)
    od, rd = O.l1_forward(of, ef, rel_inds[:, 1:3], p, T)
    cases.check_rows(od, fx['obj_rows'], fx['obj_dists'], fx['obj_colsum'], 5e-5, 'obj_dists')
    cases.check_rows(rd, fx['rel_rows'], fx['rel_dists'], fx['rel_colsum'], 5e-5, 'rel_dists')


def test_draw_union_boxes_bit_exact():
    fx = cases.load('draw_union_boxes')
    out = O.draw_union_boxes(fx['pairs'], 27)
    assert out.dtype == np.float32 and np.array_equal(out, fx['out'])


def test_union_geom_eval_and_train_bn():
    fx = cases.load('union_geom')
    rois, ui, p = cases.geom_inputs(fx)
    pools = np.zeros((ui.shape[0], 512, 7, 7), np.float32)
    ev = O.union_boxes_and_feats(pools, rois, ui, p, training=False)[:, :, 0, 0]
    tr = O.union_boxes_and_feats(pools, rois, ui, p, training=True)[:, :, 0, 0]
    assert np.abs(ev - fx['out_eval']).max() <= 2e-5
    assert np.abs(tr - fx['out_train']).max() <= 1e-4   # BN batch statistics amplify rounding


def test_roi_align_node_and_union():
    fx = cases.load('roi_align')
    fmap, rois, ui = cases.roi_inputs(fx)
    nf, ef = O.node_edge_features(fmap, rois, ui)
    assert np.abs(nf - fx['node_feat']).max() <= 1e-5
    assert np.abs(ef - fx['edge_feat']).max() <= 1e-5


def test_l2_predict():
    fx = cases.load('l2_predict')
    nfe, efe, rel_inds, rois, p = cases.l2_inputs(fx)
    od, rd = O.predict(nfe, efe, rel_inds, rois, p)
    assert np.abs(od - fx['obj_dists']).max() <= 5e-5
    assert np.abs(rd - fx['rel_dists']).max() <= 5e-5


def test_get_rel_inds_eval_order():
    im = np.array([0, 0, 0, 1, 1])
    r = O.get_rel_inds_eval(im)
    assert r.shape == (3 * 2 + 2, 3)
    assert r[0].tolist() == [0, 0, 1] and r[-1].tolist() == [1, 4, 3]


@pytest.mark.parametrize('name', ['l1_cfg1', 'l1_cfg2_s3'])
def test_torch_cpu_port_l1(name):
    import torch
    from oracle.imp_torch_cpu import ImpCpu
    fx = cases.load(name)
    of, ef, rel_inds, p, T = cases.l1_inputs(fx)
    m = ImpCpu(mp_iter=T).load_numpy(p).eval()
    with torch.no_grad():
        od, rd = m.l1_forward(torch.from_numpy(of), torch.from_numpy(ef), torch.from_numpy(rel_inds[:, 1:3]))
    cases.check_rows(od.numpy(), fx['obj_rows'], fx['obj_dists'], fx['obj_colsum'], 2e-5, 'obj_dists')
    cases.check_rows(rd.numpy(), fx['rel_rows'], fx['rel_dists'], fx['rel_colsum'], 2e-5, 'rel_dists')


def test_draw_union_boxes_vs_compiled_reference():
    """oracle/_ref = the reference's own Cython op built from /root/reference (oracle/build_ref.py)."""
    from oracle import build_ref
    mod = build_ref.load()
    if mod is None:
        pytest.skip('oracle/_ref not built (no /root/reference on this box)')
    rng = np.random.default_rng(5)
    xy = rng.random((200, 2, 2), dtype=np.float32) * 500
    wh = rng.random((200, 2, 2), dtype=np.float32) * 300 + 1
    pairs = np.concatenate((xy[:, 0], xy[:, 0] + wh[:, 0], xy[:, 1], xy[:, 1] + wh[:, 1]), 1).astype(np.float32)
    assert np.array_equal(O.draw_union_boxes(pairs, 27), mod.draw_union_boxes(pairs, 27))
