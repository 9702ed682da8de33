"""Accuracy report (GPU box): max |CUDA - oracle| per golden case and GEMM mode."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops
from oracle import imp_numpy as O
from tests import cases

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for mode in ('simt', 'tc'):
    ops.set_gemm_mode(mode)
    for name in ('l0_cfg1', 'l0_cfg2_s3', 'l0_ragged_t6', 'l0_special'):
        fx = cases.load(name)
        obj, rel, ri, p, T = cases.l0_inputs(fx)
        g = ops.build_graph(dev(ri)[:, 1:3], obj.shape[0])
        v, e = ops.message_pass(dev(rel), dev(obj), g, {k: dev(x) for k, x in p.items()}, T)
        vo, eo = O.message_pass(rel.astype(np.float64), obj.astype(np.float64), ri[:, 1:3], {k: x.astype(np.float64) for k, x in p.items()}, T)
        print('%-5s %-14s V err %.2e (|V| %.2f)  E err %.2e (|E| %.2f)' % (mode, name, np.abs(v.cpu().numpy() - vo).max(), np.abs(vo).max(),
              np.abs(e.cpu().numpy() - eo).max(), np.abs(eo).max()), flush=True)
    for name in ('l1_cfg1', 'l1_cfg2', 'l1_cfg2_s3'):
        fx = cases.load(name)
        of, ef, ri, p, T = cases.l1_inputs(fx)
        g = ops.build_graph(dev(ri)[:, 1:3], of.shape[0])
        od, rd = ops.l1_forward(dev(of), dev(ef), g, {k: dev(x) for k, x in p.items()}, T)
        p64 = {k: x.astype(np.float64) for k, x in p.items()}
        oo, ro = O.l1_forward(of.astype(np.float64), ef.astype(np.float64), ri[:, 1:3], p64, T)
        print('%-5s %-14s obj err %.2e (|.| %.2f)  rel err %.2e (|.| %.2f)' % (mode, name, np.abs(od.cpu().numpy() - oo).max(), np.abs(oo).max(),
              np.abs(rd.cpu().numpy() - ro).max(), np.abs(ro).max()), flush=True)
ops.set_gemm_mode('tc')
print('acc report done')
