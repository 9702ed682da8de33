"""Gradient clipping + SGD step of the training loop as multi-tensor CUDA sweeps (csrc/train.cu).

Mirrors the reference's helpers (same names, arguments and results):

* ``clip_grad_norm`` / ``grad_clip``  — lib/pytorch_misc.py:625-664, :70-73 (called at main.py:119)
* ``get_optim`` / ``update_lr``       — lib/pytorch_misc.py:130-157, :667-670 (main.py:238)

``FusedSGD`` is a ``torch.optim.Optimizer`` with ``torch.optim.SGD``'s state layout (``momentum_buffer``), so the
reference's ``optimizer.load_state_dict(ckpt['optimizer'])`` / ``MultiStepLR`` / ``save_checkpoint`` work unchanged.
One ``step()`` is ONE kernel over all parameter tensors (two with clipping: squared-norm partials, then the
clip + weight-decay + momentum + update sweep); in the 3xFP16 tensor-core mode the same sweep also rewrites the fp16
operand split of every weight the forward uses on tensor cores, so the next forward starts without a re-split pass.
No CPU path: CPU parameters raise.
"""
import ctypes as C
import torch

from . import _lib, ops
from ._lib import MtTensor, check
from .ops import _ptr, _stream


class _Table(object):
    """Device copy of an ``sgg_mt_tensor`` array; re-uploaded only when a pointer / lr / wd / flag changes."""

    def __init__(self):
        self.sig = None
        self.n = 0
        self.chunks = 0
        self.dev = None
        self.ws = None
        self.norm = None

    def sync(self, rows, device):
        """rows: list of (p_ptr, g_ptr, m_ptr, split_ptr, numel, lr, wd, flags)."""
        sig = tuple(rows)
        if sig == self.sig:
            return
        lib = _lib.load()
        n = len(rows)
        host = (MtTensor * n)()
        for i, (p, g, m, sp, numel, lr, wd, fl) in enumerate(rows):
            host[i].p, host[i].g, host[i].m, host[i].split = p or None, g or None, m or None, sp or None
            host[i].n, host[i].lr, host[i].wd, host[i].flags = numel, lr, wd, fl
        nbytes = lib.sgg_mt_table_bytes(n)
        if self.dev is None or self.dev.numel() < nbytes:
            self.dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
        chunks = lib.sgg_mt_total_chunks(host, n)
        wsb = lib.sgg_mt_workspace_bytes(chunks)
        if self.ws is None or self.ws.numel() < wsb:
            self.ws = torch.empty(wsb, dtype=torch.uint8, device=device)
        if self.norm is None:
            self.norm = torch.zeros(4, dtype=torch.float32, device=device)
        check(lib.sgg_mt_table_upload(host, n, _ptr(self.dev), self.dev.numel(), _stream()), 'sgg_mt_table_upload')
        self.sig, self.n, self.chunks = sig, n, chunks


def _grad_of(p):
    g = p.grad
    if g is None:
        return None
    if not g.is_cuda or g.dtype != torch.float32:
        raise _lib.SggError('gradients must be CUDA float32 tensors (sgg_b200 has no CPU path)')
    if g.is_sparse:
        raise RuntimeError('FusedSGD does not support sparse gradients')
    if not g.is_contiguous():
        p.grad = g = g.contiguous()
    return g


_CLIP_TABLES = {}


def clip_grad_norm(named_parameters, max_norm, clip=False, verbose=False):
    """Global L2 norm over all gradients, as if concatenated (lib/pytorch_misc.py:625-664); when ``clip`` and
    ``max_norm / (norm + 1e-6) < 1`` the gradients are scaled in place.  Returns the total norm (0-d CUDA tensor; the
    reference returns the same).  Two launches whatever the number of tensors; nothing synchronises unless
    ``verbose``."""
    named_parameters = list(named_parameters)
    grads = [(n, _grad_of(p)) for n, p in named_parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    lib = _lib.load()
    dev = grads[0][1].device
    tab = _CLIP_TABLES.setdefault((dev, len(grads)), _Table())
    tab.sync([(0, g.data_ptr(), 0, 0, g.numel(), 0.0, 0.0, 0) for _, g in grads], dev)
    check(lib.sgg_mt_grad_norm(_ptr(tab.dev), tab.n, tab.chunks, float(max_norm), _ptr(tab.norm), _ptr(tab.ws),
                               tab.ws.numel(), _stream()), 'sgg_mt_grad_norm')
    if clip:
        check(lib.sgg_mt_scale_grads(_ptr(tab.dev), tab.n, tab.chunks, _ptr(tab.norm), _stream()), 'sgg_mt_scale_grads')
        for _, g in grads:
            torch.autograd.graph.increment_version(g)
    total = tab.norm[0].clone()
    if verbose:                                    # logging only (lib/pytorch_misc.py:657-662)
        coef = float(max_norm) / (float(total) + 1e-6)
        print('---Total norm {:.3f} clip coef {:.3f}-----------------'.format(float(total), coef))
        per = sorted(((n, float(torch.linalg.vector_norm(g)) * (coef if clip and coef < 1 else 1.0), tuple(g.shape))
                      for n, g in grads), key=lambda x: -x[1])
        for name, norm, shape in per:
            print('{:<50s}: {:.3f}, ({})'.format(name, norm, shape))
        print('-------------------------------', flush=True)
    return total


def grad_clip(detector, clip, verbose):
    """lib/pytorch_misc.py:70-73."""
    return clip_grad_norm([(n, p) for n, p in detector.named_parameters() if p.grad is not None],
                          max_norm=clip, verbose=verbose, clip=True)


class FusedSGD(torch.optim.Optimizer):
    """``torch.optim.SGD(params, lr, momentum, weight_decay)`` (dampening 0, no Nesterov) as one multi-tensor sweep.

    ``step(max_norm=c)`` folds the reference's ``grad_clip(model, c)`` into the step: the clip factor is computed
    on the device and applied to the gradient inside the update (``write_clipped_grads=True`` also stores the
    scaled gradient back, reproducing the in-place side effect of the reference's clip)."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, emit_operand_split=True):
        if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0:
            raise ValueError('invalid lr / momentum / weight_decay')
        defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0, nesterov=False,
                        maximize=False, foreach=None, differentiable=False, fused=None)
        super().__init__(params, defaults)
        moms = {g['momentum'] for g in self.param_groups}
        if len(moms) > 1:
            raise ValueError('FusedSGD: one momentum for all groups')
        self.emit_operand_split = emit_operand_split
        self._table = _Table()
        self.last_norm = None

    @torch.no_grad()
    def step(self, closure=None, max_norm=None, write_clipped_grads=False, grad_scale=1.0):
        """``grad_scale``: the gradients in memory are the data-parallel SUM; the step uses grad_scale * g
        (``FlatGradReducer.grad_scale`` = 1 / world), folded into the clip factor — no separate averaging sweep."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        tc16 = self.emit_operand_split and ops._use_tc() and lib.sgg_tc_get_mode() == 1
        rows, touched, split_ps, device = [], [], [], None
        momentum = 0.0
        for group in self.param_groups:
            momentum = float(group['momentum'])
            for p in group['params']:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.SggError('FusedSGD parameters must be contiguous CUDA float32 tensors')
                device = p.device
                g = _grad_of(p)
                st = self.state[p]
                buf = st.get('momentum_buffer')
                first = buf is None
                if first:
                    if g is None:
                        continue                   # nothing to do yet; the buffer is created with the first gradient
                    buf = st['momentum_buffer'] = torch.empty_like(p, memory_format=torch.contiguous_format)
                sp = None
                if tc16 and g is not None and p.dim() == 2 and p.numel() % 8 == 0:
                    sp = ops.split_cache_peek(p)
                    if sp is not None:
                        split_ps.append(p)
                rows.append((p.data_ptr(), g.data_ptr() if g is not None else 0, buf.data_ptr(),
                             sp.data_ptr() if sp is not None else 0, p.numel(), float(group['lr']),
                             float(group['weight_decay']), 1 if first else 0))
                if g is not None:
                    touched.append((p, g))
        if not rows:
            return loss
        tab = self._table
        tab.sync(rows, device)
        norm_ptr = C.c_void_p(0)
        if max_norm is not None:
            check(lib.sgg_mt_grad_norm_scaled(_ptr(tab.dev), tab.n, tab.chunks, float(max_norm), float(grad_scale),
                                              _ptr(tab.norm), _ptr(tab.ws), tab.ws.numel(), _stream()),
                  'sgg_mt_grad_norm_scaled')
            norm_ptr = _ptr(tab.norm)
            self.last_norm = tab.norm
        elif grad_scale != 1.0:                    # no clipping: the sweep still applies norm[2] = grad_scale
            if getattr(self, '_scale_only', None) is None or float(self._scale_only[1]) != float(grad_scale):
                t = torch.tensor([0.0, 1.0, float(grad_scale), 0.0], dtype=torch.float32, device=device)
                self._scale_only = (t, float(grad_scale))
            norm_ptr = _ptr(self._scale_only[0])
        check(lib.sgg_mt_sgd_step(_ptr(tab.dev), tab.n, tab.chunks, norm_ptr, momentum, 1 if write_clipped_grads else 0,
                                  _stream()), 'sgg_mt_sgd_step')
        for p, g in touched:                       # raw-pointer writes: tell autograd / the split cache about them
            torch.autograd.graph.increment_version(p)
            if write_clipped_grads and max_norm is not None:
                torch.autograd.graph.increment_version(g)
        for p in split_ps:
            ops.split_cache_commit(p)
        return loss


def get_optim(detector, lr, conf, start_epoch, ckpt=None):
    """lib/pytorch_misc.py:130-157: SGD(momentum 0.9, weight_decay conf.l2) with the VGG fc layers (names starting
    with ``roi_fmap``) at lr / 10, optional optimizer-state restore, MultiStepLR(milestones = conf.steps + 1)."""
    print('\nEffective learning rate is %.3e' % lr)
    fc_params = [(n, p) for n, p in detector.named_parameters() if n.startswith('roi_fmap') and p.requires_grad]
    non_fc_params = [(n, p) for n, p in detector.named_parameters() if not n.startswith('roi_fmap') and p.requires_grad]
    params = [{'params': [p for _, p in fc_params], 'lr': lr / 10.0},
              {'params': [p for _, p in non_fc_params]}]
    optimizer = FusedSGD(params, weight_decay=conf.l2, lr=lr, momentum=0.9)
    if start_epoch > -1 and ckpt is not None:
        print('Restoring optimizers')
        try:
            optimizer.load_state_dict(ckpt['optimizer'])
        except Exception as e:                     # the reference swallows this too
            print('error restoring optimizer', e)
    scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=[s + 1 for s in conf.steps],
                                                     gamma=conf.lr_decay)
    return optimizer, scheduler


def update_lr(optimizer, lr=1e-4):
    """lib/pytorch_misc.py:667-670."""
    print('------ Learning rate -> {}'.format(lr))
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr
