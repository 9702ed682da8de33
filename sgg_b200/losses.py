"""Classification losses of the training step, mirroring ``lib/losses.py`` of the reference.

``edge_losses`` / ``node_losses`` keep the reference's names, arguments, return dictionaries and error
behaviour (lib/losses.py:5-74; called from main.py:105-114).  Each call is one C-ABI call (``sgg_ce_loss``):
row-wise log-softmax, the reference's FG/BG row weighting, the summed loss and d loss / d logits are produced in a
single pass, so ``backward`` only scales the stored gradient by the incoming scalar.  No CPU path.
"""
import torch

from . import _lib
from ._lib import check
from .ops import _f32, _ptr, _stream

_MODES = {'mean': 0, 'baseline': 1, 'dnorm': 2, 'dnorm-fgbg': 3}


def _ce(logits, labels, mode, alpha=1.0, beta=1.0, gamma=1.0, category=None, need_grad=True, validate=False):
    """-> (loss [] fp32, dlogits [M,C] or None, counts int32[4] device tensor)."""
    lib = _lib.load()
    x = _f32(logits, 'logits')
    if x.dim() != 2:
        raise _lib.SggError('logits must be [M, C]')
    M, Cn = x.shape
    if not isinstance(labels, torch.Tensor) or not labels.is_cuda or labels.dtype != torch.int64:
        raise _lib.SggError('labels must be a CUDA int64 tensor')
    lab = labels.detach().contiguous().view(-1)
    assert M == lab.numel(), (M, lab.numel())                      # lib/losses.py:34
    if category is not None:
        category = category.contiguous()
        assert category.dtype == torch.int8 and category.numel() == M
    ws = torch.empty(lib.sgg_ce_loss_workspace_bytes(M), dtype=torch.uint8, device=x.device)
    loss = torch.empty((), dtype=torch.float32, device=x.device)
    counts = torch.empty(4, dtype=torch.int32, device=x.device)
    dl = torch.empty_like(x) if need_grad else None
    check(lib.sgg_ce_loss(_ptr(x), _ptr(lab), _ptr(category), M, Cn, _MODES[mode], float(alpha), float(beta),
                          float(gamma), _ptr(loss), _ptr(dl), _ptr(counts), _ptr(ws), ws.numel(), _stream()),
          'sgg_ce_loss')
    if validate and int(counts[3]) != 0:
        raise IndexError('Target out of bounds: %d labels outside [0, %d)' % (int(counts[3]), Cn))
    return loss, dl, counts


class _CeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, mode, alpha, beta, gamma, category):
        loss, dl, counts = _ce(logits, labels, mode, alpha, beta, gamma, category,
                               need_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(dl)
        ctx.mark_non_differentiable(counts)
        return loss, counts

    @staticmethod
    def backward(ctx, dloss, _dcounts):
        (dl,) = ctx.saved_tensors
        return (dl * dloss if dl is not None else None), None, None, None, None, None, None


def cross_entropy_sum(logits, labels, mode='mean', alpha=1.0, beta=1.0, gamma=1.0, category=None):
    """Weighted cross entropy (already reduced the way ``mode`` says) + the device-side {M_FG, M_BG, ...} counts."""
    return _CeLossFn.apply(logits, labels, mode, alpha, beta, gamma, category)


def edge_losses(rel_dists, rel_labels, loss_type='dnorm', idx_fg=None, idx_bg=None, return_idx=False,
                loss_weights=(1, 1, 1), sfx=''):
    """Predicate classification loss (lib/losses.py:5-70): 'baseline' = gamma * mean CE over all edges;
    'dnorm' = FG edges weighted alpha / M_FG and BG edges beta / M_FG; 'dnorm-fgbg' = BG edges beta / M_BG.
    Returns ``{'rel_loss' + sfx: loss}`` (and idx_fg, idx_bg when ``return_idx``)."""
    alpha, beta, gamma = loss_weights
    if loss_type == 'baseline':
        assert alpha == beta == 1, ('wrong loss is used, use dnorm or dnorm-fgbg', alpha, beta)   # lib/losses.py:42
    elif loss_type not in ('dnorm', 'dnorm-fgbg'):
        raise NotImplementedError(loss_type)                                                       # lib/losses.py:66
    assert len(rel_dists) == len(rel_labels), (len(rel_dists), len(rel_labels))
    category = None
    if idx_fg is not None or idx_bg is not None:
        # explicit index sets (reused by the caller for a second batch, main.py:164-169): rows listed in neither
        # keep weight 1, exactly like the reference's edge_weights = ones(M)
        if idx_fg is None:
            idx_fg = torch.nonzero(rel_labels > 0).view(-1)
        if idx_bg is None:
            idx_bg = torch.nonzero(rel_labels == 0).view(-1)
        category = torch.zeros(len(rel_labels), dtype=torch.int8, device=rel_dists.device)
        category[idx_fg] = 1
        category[idx_bg] = 2
    loss, _ = cross_entropy_sum(rel_dists, rel_labels, loss_type, alpha, beta, gamma, category)
    losses = {'rel_loss' + sfx: loss}
    if return_idx:
        if idx_fg is None:
            idx_fg = torch.nonzero(rel_labels > 0).view(-1)
        if idx_bg is None:
            idx_bg = torch.nonzero(rel_labels == 0).view(-1)
        return losses, idx_fg, idx_bg
    return losses


def node_losses(rm_obj_dists, rm_obj_labels, sfx=''):
    """Object classification loss (lib/losses.py:73-74): mean cross entropy."""
    loss, _ = cross_entropy_sum(rm_obj_dists, rm_obj_labels, 'mean')
    return {'obj_loss' + sfx: loss}
