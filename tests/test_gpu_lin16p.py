"""GPU: the pre-split LINEAR kernel (csrc/lin16p.cu, `sgg_tc16_linear_pre`) and the operand planes RoIAlign emits for it
(`sgg_node_edge_features_planes`) — the eval-mode fc6 / fc7 path of rel_model_stanford.py:100-101.  Oracle: float64 matmul
of the same fp32 inputs (bar 1e-5 relative to the output scale: fp32-grade), and the fp32-input engine."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _planes(x):
    hi = x.half()
    lo = ((x - hi.float()) * 2048.0).half()
    return torch.stack((hi, lo)).contiguous()


@pytest.mark.parametrize('shape', [(300, 132, 520), (128, 128, 64), (1000, 4096, 1024), (37, 8, 8)])
@pytest.mark.parametrize('relu', [False, True])
def test_linear_pre_split_vs_float64(shape, relu):
    from sgg_b200 import ops
    if ops.tc_engine() != 'tc16':
        pytest.skip('3xFP16 engine not selected')
    M, Nout, K = shape
    g = torch.Generator(device='cuda').manual_seed(M + Nout + K)
    x = torch.randn(M, K, device='cuda', generator=g) * 3.0
    w = torch.nn.Parameter(torch.randn(Nout, K, device='cuda', generator=g) / K ** 0.5)
    b = torch.randn(Nout, device='cuda', generator=g)
    ref = x.double() @ w.detach().double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    y, ypl = ops.linear(x, w, b, relu=relu, x_planes=_planes(x), out_planes=True)
    scale = float(ref.abs().max())
    assert float((y.double() - ref).abs().max()) <= 1e-5 * scale
    y2 = ops.linear(x, w, b, relu=relu)                       # fp32-input engine, same weights
    assert float((y - y2).abs().max()) <= 2e-5 * scale
    assert ypl is not None and ypl.shape == (2, M, Nout)
    exp = _planes(y)
    assert torch.equal(ypl[0], exp[0]) and torch.equal(ypl[1], exp[1])       # the emitted planes ARE the split of y
    assert ops.tc16_overflow(reset=True) == 0


def test_linear_pre_split_chain_and_bad_planes():
    from sgg_b200 import ops, _lib
    if ops.tc_engine() != 'tc16':
        pytest.skip('3xFP16 engine not selected')
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn(260, 512, device='cuda', generator=g)
    w1 = torch.nn.Parameter(torch.randn(256, 512, device='cuda', generator=g) / 22.6)
    w2 = torch.nn.Parameter(torch.randn(64, 256, device='cuda', generator=g) / 16.0)
    h, hpl = ops.linear(x, w1, None, relu=True, x_planes=_planes(x), out_planes=True)
    y = ops.linear(h, w2, None, x_planes=hpl)                 # second layer consumes the planes of the first
    ref = (x.double() @ w1.detach().double().t()).clamp_min(0) @ w2.detach().double().t()
    assert float((y.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    with pytest.raises(_lib.SggError):
        ops.linear(x, w1, None, x_planes=_planes(x)[:, :100])                # wrong shape: loud, no silent fallback


def test_roi_align_planes_are_the_split_of_the_fp32_rows():
    from sgg_b200 import ops, synth
    gph = synth.synth_graph(2, 12, 40, 3)
    rois = torch.from_numpy(gph['rois']).cuda()
    rel = torch.from_numpy(gph['rel_inds']).cuda()
    fmap = torch.rand(2, 64, 38, 38, device='cuda', generator=torch.Generator(device='cuda').manual_seed(2)) * 4
    add = torch.randn(rel.shape[0], 64, device='cuda', generator=torch.Generator(device='cuda').manual_seed(3))
    n0, e0 = ops.node_edge_features(fmap, rois, rel[:, 1:3], edge_add=add)
    n1, e1, npl, epl = ops.node_edge_features(fmap, rois, rel[:, 1:3], edge_add=add, planes=True)
    assert torch.equal(n0, n1) and torch.equal(e0, e1)
    for rows, pl in ((n1, npl), (e1, epl)):
        exp = _planes(rows.reshape(rows.shape[0], -1))
        assert pl.shape == exp.shape and torch.equal(pl[0], exp[0]) and torch.equal(pl[1], exp[1])


def test_planes_only_path_matches_the_fp32_path():
    """node_edge_features(planes='only') writes no fp32 rows; linear(None, ..., x_planes=...) consumes the planes."""
    from sgg_b200 import ops, synth
    if ops.tc_engine() != 'tc16':
        pytest.skip('3xFP16 engine not selected')
    gph = synth.synth_graph(2, 12, 40, 3)
    rois = torch.from_numpy(gph['rois']).cuda()
    rel = torch.from_numpy(gph['rel_inds']).cuda()
    fmap = torch.rand(2, 64, 38, 38, device='cuda', generator=torch.Generator(device='cuda').manual_seed(2)) * 4
    n1, e1, npl, epl = ops.node_edge_features(fmap, rois, rel[:, 1:3], planes=True)
    n2, e2, npl2, epl2 = ops.node_edge_features(fmap, rois, rel[:, 1:3], planes='only')
    assert n2 is None and e2 is None and torch.equal(npl, npl2) and torch.equal(epl, epl2)
    w = torch.nn.Parameter(torch.randn(96, 64 * 49, device='cuda', generator=torch.Generator(device='cuda').manual_seed(4)) / 56.0)
    ya = ops.linear(e1.reshape(e1.shape[0], -1), w, None, relu=True, x_planes=epl)
    yb = ops.linear(None, w, None, relu=True, x_planes=epl2)
    assert torch.equal(ya, yb)
    ref = (e1.reshape(e1.shape[0], -1).double() @ w.detach().double().t()).clamp_min(0)
    assert float((yb.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_bcast_add_planes_are_the_split_of_the_sum():
    from sgg_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(9)
    pools = torch.rand(37, 16, 7, 7, device='cuda', generator=g) * 5
    geom = torch.randn(37, 16, device='cuda', generator=g)
    out0 = ops.bcast_add(pools, geom)
    out, pl = ops.bcast_add(pools, geom, planes=True)
    assert torch.equal(out, out0) and torch.equal(out, pools + geom[:, :, None, None])
    exp = _planes(out.reshape(37, -1))
    assert pl.shape == exp.shape and torch.equal(pl[0], exp[0]) and torch.equal(pl[1], exp[1])
