"""GPU parity of the training tail (csrc/train.cu through the C-ABI): sgg_b200.losses / sgg_b200.optim against
 (a) golden vectors produced by running the reference (lib/losses.py, clip_grad_norm, get_optim's SGD),
 (b) the numpy oracle, and (c) torch.optim.SGD on the device at a size the fixtures cannot hold.
Bars: loss 1e-5 relative; d loss / d logits 1e-5 of its scale; parameters / momentum 2e-6 relative (fp32, same
operation sequence as torch)."""
import types
import numpy as np
import pytest
import torch

from oracle import imp_numpy as O
from sgg_b200 import synth
from tests import cases

pytestmark = pytest.mark.gpu
FX = cases.load('train_tail')


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('name', sorted(synth.loss_cases()))
def test_losses_vs_golden_and_oracle(name):
    from sgg_b200 import losses
    c = synth.loss_cases()[name]
    logits, labels = synth.synth_logits(c['M'], c['C'], c['seed'], fg=c['fg'])
    x = dev(logits).requires_grad_(True)
    y = dev(labels)
    if c['kind'] == 'node':
        out = losses.node_losses(x, y)
        assert list(out) == ['obj_loss']
        loss = out['obj_loss']
        oloss, og = O.node_losses(logits, labels, return_grad=True)
    else:
        kw, okw = {}, {}
        if c.get('explicit_idx'):
            fg, bg = synth.explicit_idx(labels, c['seed'])
            kw = dict(idx_fg=dev(fg), idx_bg=dev(bg)); okw = dict(idx_fg=fg, idx_bg=bg)
        out = losses.edge_losses(x, y, c['kind'], loss_weights=c['w'], sfx='_t', **kw)
        assert list(out) == ['rel_loss_t']
        loss = out['rel_loss_t']
        oloss, og = O.edge_losses(logits, labels, c['kind'], c['w'], return_grad=True, **okw)
    (2.0 * loss).backward()                     # upstream scalar gradient must be honoured
    ref = float(FX['loss_' + name])
    assert abs(float(loss) - ref) <= 1e-5 * max(1.0, abs(ref)), (float(loss), ref)
    assert abs(float(loss) - oloss) <= 1e-5 * max(1.0, abs(oloss))
    g = x.grad.cpu().numpy() / 2.0
    gs = max(np.abs(og).max(), 1e-12)
    assert np.abs(g - og).max() <= 1e-5 * gs
    assert np.abs(g[FX['dlogits_rows_' + name]] - FX['dlogits_' + name]).max() <= 1e-5 * gs


def test_edge_losses_return_idx_and_errors():
    from sgg_b200 import losses
    logits, labels = synth.synth_logits(300, 51, 3, fg=0.1)
    x, y = dev(logits), dev(labels)
    out, fg, bg = losses.edge_losses(x, y, 'dnorm', return_idx=True)
    assert np.array_equal(fg.cpu().numpy(), np.nonzero(labels > 0)[0])
    assert np.array_equal(bg.cpu().numpy(), np.nonzero(labels == 0)[0])
    assert abs(float(out['rel_loss']) - O.edge_losses(logits, labels, 'dnorm')) <= 1e-5
    with pytest.raises(AssertionError):
        losses.edge_losses(x, y, 'baseline', loss_weights=(2, 1, 1))
    with pytest.raises(NotImplementedError):
        losses.edge_losses(x, y, 'focal')
    with pytest.raises(AssertionError):
        losses.edge_losses(x, y[:-1], 'dnorm')


def test_ignore_index_and_full_size_loss_property():
    """cfg4 per-GPU size (E = 9600): the loss of a row-permuted batch is the same number, rows labelled -100 do not
    count for the mean and get zero gradient (F.cross_entropy's default ignore_index)."""
    from sgg_b200 import losses
    logits, labels = synth.synth_logits(9600, 51, 21, fg=0.05)
    x, y = dev(logits), dev(labels)
    a = float(losses.edge_losses(x, y, 'dnorm-fgbg')['rel_loss'])
    perm = torch.randperm(9600, generator=torch.Generator().manual_seed(1)).cuda()
    b = float(losses.edge_losses(x[perm].contiguous(), y[perm].contiguous(), 'dnorm-fgbg')['rel_loss'])
    assert abs(a - b) <= 2e-6 * abs(a)
    y2 = y.clone(); y2[::3] = -100
    xr = x.clone().requires_grad_(True)
    l2 = losses.node_losses(xr, y2)['obj_loss']
    l2.backward()
    ref = torch.nn.functional.cross_entropy(x, y2)
    assert abs(float(l2) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float(xr.grad[::3].abs().max()) == 0.0


class _Net(torch.nn.Module):
    def __init__(self, params):
        super().__init__()
        self.names = list(params)
        for n, v in params.items():
            self.register_parameter(n.replace('.', '_'), torch.nn.Parameter(dev(v.copy())))

    def named_parameters(self, *a, **k):        # reference-style dotted names (get_optim keys on the 'roi_fmap' prefix)
        for n in self.names:
            yield n, getattr(self, n.replace('.', '_'))


@pytest.mark.parametrize('flow', ['reference', 'fused'])
def test_clip_and_sgd_vs_golden(flow):
    """reference flow = grad_clip(model, clip) then optimizer.step() (main.py:119-120); fused flow =
    optimizer.step(max_norm=clip).  Both must land on the parameters the reference produced."""
    from sgg_b200 import optim
    tt = synth.synth_train_tail(seed=5)
    net = _Net(tt['params'])
    conf = types.SimpleNamespace(l2=tt['l2'], steps=[15], lr_decay=0.1)
    opt, sched = optim.get_optim(net, tt['lr'], conf, -1)
    assert sorted(g['lr'] for g in opt.param_groups) == sorted(FX['group_lrs'].tolist())
    for step in range(tt['steps']):
        for n, p in net.named_parameters():
            g = tt['grads'][step][n]
            p.grad = None if g is None else dev(g.copy())
        if flow == 'reference':
            tn = optim.grad_clip(net, tt['clip'], verbose=(step == 0))
            if step == 0:
                for n, p in net.named_parameters():
                    np.testing.assert_allclose(p.grad.cpu().numpy(), FX['clipped_grad0_' + n], rtol=2e-6, atol=1e-9)
            opt.step()
        else:
            opt.step(max_norm=tt['clip'])
            tn = opt.last_norm[0]
        assert abs(float(tn) - FX['norms'][step]) <= 2e-6 * FX['norms'][step]
        if step in (0, tt['steps'] - 1):
            for n, p in net.named_parameters():
                np.testing.assert_allclose(p.detach().cpu().numpy(), FX['p_step%d_%s' % (step, n)], rtol=2e-6, atol=2e-8,
                                           err_msg='%s step %d' % (n, step))
    for n, p in net.named_parameters():
        np.testing.assert_allclose(opt.state[p]['momentum_buffer'].cpu().numpy(), FX['m_final_' + n], rtol=2e-6,
                                   atol=2e-8, err_msg=n)
    # state layout is torch.optim.SGD's: a reference checkpoint's optimizer state loads, and ours loads into torch's
    ref_opt = torch.optim.SGD([{'params': g['params'], 'lr': g['lr']} for g in opt.param_groups], lr=tt['lr'],
                              momentum=0.9, weight_decay=tt['l2'])
    ref_opt.load_state_dict(opt.state_dict())
    opt.load_state_dict(ref_opt.state_dict())


def test_fused_sgd_matches_torch_sgd_at_scale():
    """~60 M parameters in the shapes of the model's big tensors (fc7 4096x4096, the unary 512x4096 pair, GRU
    matrices, biases, a ragged tail), 3 steps with clipping engaged on the first: device-side comparison with
    torch.optim.SGD + the reference's clip formula."""
    from sgg_b200 import optim
    gen = torch.Generator(device='cuda').manual_seed(3)
    shapes = [(4096, 4096), (4096, 4096), (512, 4096), (512, 4096), (1536, 512), (1536, 512), (4096,), (1536,),
              (151, 512), (51, 512), (1, 1024), (1,), (256, 2, 7, 7), (3, 1365)]
    ours = [torch.nn.Parameter(0.05 * torch.randn(s, device='cuda', generator=gen)) for s in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    o1 = optim.FusedSGD([{'params': ours[:2], 'lr': 0.01}, {'params': ours[2:]}], lr=0.1, momentum=0.9, weight_decay=1e-4)
    o2 = torch.optim.SGD([{'params': theirs[:2], 'lr': 0.01}, {'params': theirs[2:]}], lr=0.1, momentum=0.9, weight_decay=1e-4)
    for step, gs in enumerate((3e-3, 1e-4, 1e-4)):
        for a, b in zip(ours, theirs):
            g = gs * torch.randn(a.shape, device='cuda', generator=gen)
            a.grad, b.grad = g.clone(), g.clone()
        total = torch.sqrt(sum((b.grad.double() ** 2).sum() for b in theirs))
        coef = 5.0 / (float(total) + 1e-6)
        if coef < 1:
            for b in theirs:
                b.grad.mul_(coef)
        assert (coef < 1) == (step == 0)
        o1.step(max_norm=5.0)
        o2.step()
        assert abs(float(o1.last_norm[0]) - float(total)) <= 2e-6 * float(total)
        for a, b in zip(ours, theirs):
            d = float((a.detach() - b.detach()).abs().max())
            assert d <= 2e-7, (step, tuple(a.shape), d)        # |p| ~ 0.05..0.25: one or two ulps
            dm = float((o1.state[a]['momentum_buffer'] - o2.state[b]['momentum_buffer']).abs().max())
            assert dm <= 2e-8 + 2e-6 * float(o2.state[b]['momentum_buffer'].abs().max())


def test_fused_sgd_rewrites_the_operand_split():
    """3xFP16 mode: the step also emits the fp16 [hi | lo] split of each weight the forward has used on tensor cores;
    the next forward must find it (no re-split launch) and it must be bit-identical to a fresh split."""
    from sgg_b200 import ops, optim, _lib
    ops.set_gemm_mode('tc16')
    try:
        lib = _lib.load()
        gen = torch.Generator(device='cuda').manual_seed(5)
        w = torch.nn.Parameter(0.1 * torch.randn(512, 4096, device='cuda', generator=gen))
        b = torch.nn.Parameter(torch.zeros(512, device='cuda'))
        x = torch.randn(300, 4096, device='cuda', generator=gen)
        y0 = ops.linear(x, w, b)                           # first use: split is computed and cached
        opt = optim.FusedSGD([w, b], lr=0.5, momentum=0.9)
        w.grad = 0.01 * torch.randn(w.shape, device='cuda', generator=gen)
        b.grad = torch.ones_like(b)
        opt.step()
        n0 = lib.sgg_launch_count()
        sp = ops.split_weight(w)                           # must be a cache hit: nothing launches
        assert lib.sgg_launch_count() == n0
        emitted = sp.clone()
        ops._SPLIT_CACHE.clear()
        fresh = ops.split_weight(w)
        n_half = w.numel()                                 # the fp16 pair occupies the first half of the buffer
        assert torch.equal(emitted.view(-1)[:n_half], fresh.view(-1)[:n_half])
        y1 = ops.linear(x, w, b)
        ref = torch.nn.functional.linear(x.double(), w.detach().double(), b.detach().double()).float()
        assert float((y1 - ref).abs().max()) <= 1e-4 and float((y1 - y0).abs().max()) > 1e-3
    finally:
        ops.set_gemm_mode('tc')


def test_clip_grad_norm_standalone_and_noop():
    from sgg_b200 import optim
    gen = torch.Generator(device='cuda').manual_seed(7)
    ps = [torch.nn.Parameter(torch.zeros(s, device='cuda')) for s in ((1000, 37), (5,), (4096, 8))]
    for p in ps:
        p.grad = torch.randn(p.shape, device='cuda', generator=gen)
    ps.append(torch.nn.Parameter(torch.zeros(3, device='cuda')))          # no gradient: ignored
    named = [('p%d' % i, p) for i, p in enumerate(ps)]
    before = [p.grad.clone() for p in ps[:3]]
    ref = float(torch.sqrt(sum((g.double() ** 2).sum() for g in before)))
    tn = optim.clip_grad_norm(named, max_norm=1e9, clip=True)             # coef >= 1: gradients untouched
    assert abs(float(tn) - ref) <= 2e-6 * ref
    assert all(torch.equal(p.grad, g) for p, g in zip(ps, before))
    tn = optim.clip_grad_norm(named, max_norm=1.0, clip=False)            # clip=False (the default): norm only
    assert all(torch.equal(p.grad, g) for p, g in zip(ps, before))
    tn = optim.clip_grad_norm(named, max_norm=1.0, clip=True)
    coef = 1.0 / (ref + 1e-6)
    for p, g in zip(ps, before):
        assert float((p.grad - g * coef).abs().max()) <= 1e-6 * coef * float(g.abs().max())
    assert float(optim.clip_grad_norm([], 1.0)) == 0.0
    with pytest.raises(Exception):
        cpu = torch.nn.Parameter(torch.zeros(3)); cpu.grad = torch.ones(3)
        optim.clip_grad_norm([('c', cpu)], 1.0)
