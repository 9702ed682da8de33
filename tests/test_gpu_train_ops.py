"""GPU: the CUDA pieces of the training path added in round 2 (csrc/bn.cu, tensor-core linear backward, the fc6 broadcast
shortcut) against torch autograd on the same device (float64 where the comparison needs head-room)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max()) / max(1.0, float(b.double().abs().max()))


def test_bn_train_forward_backward_vs_torch():
    from sgg_b200 import autograd as K
    g = torch.Generator(device='cuda').manual_seed(0)
    for M, C in [(480, 256), (120, 512), (38400, 256), (7, 64)]:
        x = torch.randn(M, C, device='cuda', generator=g) * 2 + 0.3
        bn = torch.nn.BatchNorm1d(C, momentum=0.01).cuda().train()
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.uniform_(-0.5, 0.5, generator=g)
            bn.running_mean.uniform_(-1, 1, generator=g); bn.running_var.uniform_(0.5, 2, generator=g)
        ref_bn = torch.nn.BatchNorm1d(C, momentum=0.01).cuda().train().double()
        ref_bn.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
        xa = x.clone().requires_grad_(); xb = x.double().clone().requires_grad_()
        r = torch.randn(M, C, device='cuda', generator=g)
        ya = K.relu_bn_train(xa, bn)
        yb = ref_bn(F.relu(xb))
        (ya * r).sum().backward(); (yb * r.double()).sum().backward()
        assert _rel(ya, yb) <= 2e-6 * 10
        assert _rel(xa.grad, xb.grad) <= 1e-5
        assert _rel(bn.weight.grad, ref_bn.weight.grad) <= 1e-5 and _rel(bn.bias.grad, ref_bn.bias.grad) <= 1e-5
        assert _rel(bn.running_mean, ref_bn.running_mean) <= 1e-6 and _rel(bn.running_var, ref_bn.running_var) <= 1e-6
        assert int(bn.num_batches_tracked) == 1


def test_max4_matches_maxpool2d():
    from sgg_b200 import autograd as K
    g = torch.Generator(device='cuda').manual_seed(1)
    E, C = 333, 256
    x = torch.randn(E, 4, C, device='cuda', generator=g)
    x[5, 1] = x[5, 0]                                      # exact ties: the first maximum takes the gradient
    xa = x.clone().requires_grad_(); xb = x.clone().requires_grad_()
    ya = K._Max4Fn.apply(xa)
    # reference layout: [E, C, 2, 2] -> MaxPool2d(3, 2, 1) -> [E, C, 1, 1]
    yb = F.max_pool2d(xb.permute(0, 2, 1).reshape(E, C, 2, 2), 3, 2, 1).view(E, C)
    r = torch.randn(E, C, device='cuda', generator=g)
    (ya * r).sum().backward(); (yb * r).sum().backward()
    assert torch.equal(ya, yb) and torch.equal(xa.grad, xb.grad)


def test_elementwise_helpers():
    from sgg_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(2)
    pools = torch.randn(37, 16, 7, 7, device='cuda', generator=g); geom = torch.randn(37, 16, device='cuda', generator=g)
    # 37*16*49 is a multiple of 4
    assert torch.equal(ops.bcast_add(pools, geom), pools + geom[:, :, None, None])
    y = torch.randn(64, 52, device='cuda', generator=g); dy = torch.randn(64, 52, device='cuda', generator=g)
    assert torch.equal(ops.relu_backward(dy, y), dy * (y > 0))
    w = torch.randn(40, 16, 49, device='cuda', generator=g)
    assert _rel(ops.group_sum(w, 49), w.double().sum(2)) <= 1e-6
    a = torch.randn(300, 128, device='cuda', generator=g); b = torch.randn(128, 96, device='cuda', generator=g)
    assert _rel(ops.matmul_nn(a, b), a.double() @ b.double()) <= 1e-5        # tensor-core route (M >= 256)
    assert _rel(ops.matmul_nn(a[:50].contiguous(), b), a[:50].double() @ b.double()) <= 1e-5     # SIMT route


@pytest.mark.parametrize('E', [64, 600])
def test_fc_broadcast_shortcut_matches_generic_composition(E):
    """relu(fc6(pools + broadcast(geom))): forward identical to the generic ops, dgeom through the 7x7-summed weight equals
    the 7x7 sum of the generic dX, dW / db identical (E = 600 takes the tensor-core routes)."""
    from sgg_b200 import autograd as K
    g = torch.Generator(device='cuda').manual_seed(3)
    C, S, Nout = 32, 49, 128
    pools = torch.relu(torch.randn(E, C, 7, 7, device='cuda', generator=g))
    w = (torch.randn(Nout, C * S, device='cuda', generator=g) / (C * S) ** 0.5)
    b = torch.randn(Nout, device='cuda', generator=g) * 0.1
    r = torch.randn(E, Nout, device='cuda', generator=g)
    outs = []
    for fused in (True, False):
        geom = torch.randn(E, C, device='cuda', generator=torch.Generator(device='cuda').manual_seed(4)).requires_grad_()
        wp, bp = w.clone().requires_grad_(), b.clone().requires_grad_()
        if fused:
            y = K.fc_broadcast(pools, geom, wp, bp)
        else:
            y = K.linear(K.broadcast_add(pools, geom).reshape(E, -1), wp, bp, relu=True)
        (y * r).sum().backward()
        outs.append((y.detach(), geom.grad, wp.grad, bp.grad))
    # float64 reference for the geometry gradient
    x64 = (pools.double() + outs[0][1].new_zeros(1).double())  # placeholder to keep shapes explicit
    geom64 = torch.randn(E, C, device='cuda', generator=torch.Generator(device='cuda').manual_seed(4)).double().requires_grad_()
    y64 = torch.relu((pools.double() + geom64[:, :, None, None]).reshape(E, -1) @ w.double().t() + b.double())
    (y64 * r.double()).sum().backward()
    assert torch.equal(outs[0][0], outs[1][0])
    for a, c in zip(outs[0][2:], outs[1][2:]):
        assert _rel(a, c) <= 1e-5
    assert _rel(outs[0][1], geom64.grad) <= 2e-5 and _rel(outs[1][1], geom64.grad) <= 2e-5


def test_linear_backward_writes_into_gradient_sink_without_copy():
    """A registered sink receives dW chunk by chunk (notifications in row order) and autograd adopts a view of it."""
    from sgg_b200 import ops, autograd as K
    g = torch.Generator(device='cuda').manual_seed(5)
    M, Nout, Kd = 1500, 256, 512
    x = torch.randn(M, Kd, device='cuda', generator=g)
    w = torch.nn.Parameter(torch.randn(Nout, Kd, device='cuda', generator=g) / Kd ** 0.5)
    sink = torch.full((Nout, Kd), float('nan'), device='cuda')
    seen = []
    ops.register_grad_sink(w, sink, [(0, 128), (128, 256)], lambda p, r0, r1: seen.append((r0, r1)), owner='test')
    try:
        y = K.linear(x, w)
        (y ** 2).sum().backward()
    finally:
        ops.clear_grad_sinks(owner='test')
    assert seen == [(0, 128), (128, 256)]
    # autograd normally adopts the returned alias of the sink (no copy: checked in the regular suite); under tools that change
    # object lifetimes (compute-sanitizer did) it may clone instead, which FlatGradReducer's hook handles — values must agree
    assert torch.equal(w.grad, sink)
    ref = (2 * (x.double() @ w.detach().double().t())).t() @ x.double()
    assert _rel(sink, ref) <= 1e-5


@pytest.mark.parametrize('engine', ['tc16', 'tc32'])
@pytest.mark.parametrize('gscale', [1.0, 1e-7, 3e4])
def test_tensor_core_linear_backward_any_gradient_magnitude(engine, gscale):
    """Backward GEMMs on tensor cores: the 3xFP16 engine rescales the gradient operand by an exact power of two taken from
    its device-side absolute maximum, so gradients of ANY magnitude (1e-7: far below fp16's normal range) keep fp32-grade
    accuracy; the 3xTF32 engine needs no scale.  Compared with float64."""
    from sgg_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(7)
    old = ops._TC_BWD['engine']
    ops._TC_BWD['engine'] = engine
    try:
        for (M, N, K) in [(1500, 256, 512), (2400, 512, 4096)]:
            x = torch.randn(M, K, device='cuda', generator=g); w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
            dy = torch.randn(M, N, device='cuda', generator=g) * gscale
            dy[::7] *= 1e-3                                     # rows of very different magnitude
            dx, dw, db = ops.linear_backward(x, w, dy)
            rx, rw, rb = dy.double() @ w.double(), dy.double().t() @ x.double(), dy.double().sum(0)
            for a, b in ((dx, rx), (dw, rw), (db, rb)):
                assert float((a.double() - b).abs().max()) <= 2e-6 * float(b.abs().max()), (engine, gscale, M, N, K)
    finally:
        ops._TC_BWD['engine'] = old
