"""CPU: host-side contract of the training / evaluation tail wrappers (no kernels run): the reference's argument
checks fire before any device work, CPU tensors are refused (there is no CPU path), get_optim builds the reference's
parameter groups and FusedSGD's state is interchangeable with torch.optim.SGD's."""
import types
import pytest
import torch

from sgg_b200 import losses, optim, ops
from sgg_b200._lib import SggError


def test_edge_losses_argument_checks_match_the_reference():
    x = torch.zeros(4, 51); y = torch.zeros(4, dtype=torch.int64)
    with pytest.raises(AssertionError):                      # lib/losses.py:42
        losses.edge_losses(x, y, 'baseline', loss_weights=(1, 2, 1))
    with pytest.raises(NotImplementedError):                 # lib/losses.py:66
        losses.edge_losses(x, y, 'focal')
    with pytest.raises(AssertionError):                      # lib/losses.py:34
        losses.edge_losses(x, y[:3], 'dnorm')
    with pytest.raises(SggError):                            # no CPU path
        losses.edge_losses(x, y, 'dnorm')
    with pytest.raises(SggError):
        losses.node_losses(torch.zeros(3, 151), torch.zeros(3, dtype=torch.int64))


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.roi_fmap = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Sequential(torch.nn.Linear(8, 4)))
        self.roi_fmap_obj = torch.nn.Linear(8, 4)
        self.obj_fc = torch.nn.Linear(4, 3)
        self.frozen = torch.nn.Linear(2, 2)
        for p in self.frozen.parameters():
            p.requires_grad = False


def test_get_optim_groups_scheduler_and_state_layout(capsys):
    net = _Net()
    conf = types.SimpleNamespace(l2=1e-4, steps=[2, 4], lr_decay=0.1)
    opt, sched = optim.get_optim(net, 0.12, conf, -1)
    assert 'Effective learning rate' in capsys.readouterr().out
    g_fc, g_rest = opt.param_groups
    assert g_fc['lr'] == pytest.approx(0.012) and g_rest['lr'] == pytest.approx(0.12)       # lib/pytorch_misc.py:140-141
    assert len(g_fc['params']) == 4 and len(g_rest['params']) == 2                           # frozen params excluded
    assert all(g['momentum'] == 0.9 and g['weight_decay'] == 1e-4 for g in opt.param_groups)
    assert sched.milestones == {3: 1, 5: 1}                                                  # steps + 1 (:152)
    # torch.optim.SGD state dicts load (a reference checkpoint's 'optimizer' entry), and ours load into torch's
    ref = torch.optim.SGD([{'params': g_fc['params'], 'lr': 0.012}, {'params': g_rest['params']}], lr=0.12,
                          momentum=0.9, weight_decay=1e-4)
    for p in g_rest['params']:
        p.grad = torch.ones_like(p)
    ref.step()
    opt.load_state_dict(ref.state_dict())
    assert all('momentum_buffer' in opt.state[p] for p in g_rest['params'])
    ref.load_state_dict(opt.state_dict())
    optim.update_lr(opt, 1e-4)
    assert all(g['lr'] == 1e-4 for g in opt.param_groups)
    with pytest.raises(SggError):                            # CPU parameters: refused, nothing is silently done on the host
        opt.step()
    ckpt = {'optimizer': {'state': {}, 'param_groups': []}}  # broken checkpoint: the reference swallows the error (:149)
    optim.get_optim(net, 0.1, conf, 0, ckpt)
    assert 'error restoring optimizer' in capsys.readouterr().out


def test_clip_and_rank_refuse_cpu_tensors():
    p = torch.nn.Parameter(torch.zeros(3)); p.grad = torch.ones(3)
    with pytest.raises(SggError):
        optim.clip_grad_norm([('p', p)], 1.0, clip=True)
    assert float(optim.clip_grad_norm([('q', torch.nn.Parameter(torch.zeros(2)))], 1.0)) == 0.0   # no gradients at all
    with pytest.raises(SggError):
        ops.rank_relations(torch.zeros(2, 51), torch.ones(3), torch.zeros(2, 2, dtype=torch.int64))
    with pytest.raises(ValueError):
        optim.FusedSGD([torch.nn.Parameter(torch.zeros(1))], lr=-1.0)
