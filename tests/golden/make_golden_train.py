#!/usr/bin/env python
"""Golden vectors for the training tail (losses, gradient clipping, SGD), made by RUNNING THE REFERENCE:
``lib/losses.py`` edge_losses / node_losses, ``lib/pytorch_misc.py`` clip_grad_norm and the optimizer its
``get_optim`` builds (torch.optim.SGD, lr/10 group for names starting with ``roi_fmap``).

Build container only (needs /root/reference):   python tests/golden/make_golden_train.py
Inputs come from numpy seeds (``sgg_b200.synth.synth_train_tail``); the fixture stores seeds, an input digest and
the reference's outputs.
"""
import os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from sgg_b200 import synth  # noqa: E402
from make_golden import import_reference, subsample  # noqa: E402


def main():
    import torch
    import_reference()
    from lib.losses import edge_losses, node_losses
    from lib.pytorch_misc import clip_grad_norm, get_optim
    torch.set_num_threads(1)
    out = {}
    # ---- losses -------------------------------------------------------------------------------
    cases = synth.loss_cases()
    for name, c in cases.items():
        logits, labels = synth.synth_logits(c['M'], c['C'], c['seed'], fg=c['fg'])
        x = torch.from_numpy(logits).requires_grad_(True)
        y = torch.from_numpy(labels)
        if c['kind'] == 'node':
            loss = node_losses(x, y)['obj_loss']
        else:
            kw = {}
            if c.get('explicit_idx'):
                fg, bg = synth.explicit_idx(labels, c['seed'])
                kw = dict(idx_fg=torch.from_numpy(fg), idx_bg=torch.from_numpy(bg))
            loss = edge_losses(x, y, c['kind'], loss_weights=c['w'], **kw)['rel_loss']
        loss.backward()
        out['loss_%s' % name] = np.float64(loss.item())
        d = x.grad.numpy()
        rows, sub = subsample(d, c['seed'], 32)
        out['dlogits_rows_%s' % name] = rows
        out['dlogits_%s' % name] = sub.copy()
        out['dlogits_colsum_%s' % name] = d.astype(np.float64).sum(0)
        out['digest_%s' % name] = synth.digest(logits, labels)
    # ---- clip + SGD ----------------------------------------------------------------------------
    tt = synth.synth_train_tail(seed=5)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for n, v in tt['params'].items():
                self.register_parameter(n.replace('.', '_'), torch.nn.Parameter(torch.from_numpy(v.copy())))

    m = M()
    conf = types.SimpleNamespace(l2=tt['l2'], steps=[15], lr_decay=0.1)
    opt, _ = get_optim(m, tt['lr'], conf, -1)
    names = [n.replace('.', '_') for n in tt['params']]
    norms = []
    for step in range(tt['steps']):
        for n, orig in zip(names, tt['params']):
            g = tt['grads'][step].get(orig)
            getattr(m, n).grad = None if g is None else torch.from_numpy(g.copy())
        named = [(n, p) for n, p in m.named_parameters() if p.grad is not None]
        tn = clip_grad_norm(named, max_norm=tt['clip'], clip=True)
        norms.append(float(tn))
        if step == 0:
            for n, orig in zip(names, tt['params']):
                if getattr(m, n).grad is not None:
                    out['clipped_grad0_' + orig] = getattr(m, n).grad.numpy().copy()
        opt.step()
        if step in (0, tt['steps'] - 1):
            for n, orig in zip(names, tt['params']):
                out['p_step%d_%s' % (step, orig)] = getattr(m, n).detach().numpy().copy()
    for n, orig in zip(names, tt['params']):
        st = opt.state[getattr(m, n)]
        if 'momentum_buffer' in st and st['momentum_buffer'] is not None:
            out['m_final_' + orig] = st['momentum_buffer'].numpy().copy()
    out['norms'] = np.array(norms, np.float64)
    out['digest_train_tail'] = synth.digest(*[tt['params'][k] for k in tt['params']],
                                            *[g for s in tt['grads'] for g in s.values() if g is not None])
    out['group_lrs'] = np.array([g['lr'] for g in opt.param_groups], np.float64)
    np.savez_compressed(os.path.join(HERE, 'train_tail.npz'), **out)
    print('wrote train_tail.npz: %d arrays, norms %s' % (len(out), norms))


if __name__ == '__main__':
    main()
