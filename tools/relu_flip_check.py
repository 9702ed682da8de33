"""GPU: is the edge_unary.weight gradient deviation at E = 9600 explained by ReLU-boundary flips?  Compares the CUDA
pre-activations with float64 ones and lists the (edge, unit) pairs whose sign differs."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops, synth
from tests import cases
fx = cases.load('grad_l1_shard32')
of, ef, rel_inds, p, T = cases.l1_inputs(fx)
x = torch.from_numpy(ef).cuda(); w = torch.from_numpy(p['edge_unary.weight']).cuda(); b = torch.from_numpy(p['edge_unary.bias']).cuda()
pre64 = x.double() @ w.double().t() + b.double()
for mode in ('tc16', 'simt'):
    ops.set_gemm_mode(mode)
    y = ops.linear(x, w, b, relu=True)
    flips = ((y > 0) != (pre64 > 0)).nonzero()
    print(mode, 'sign flips vs float64:', flips.shape[0], 'of', y.numel(), '| |pre64| at flips:',
          [float('%.2e' % abs(float(pre64[i, j]))) for i, j in flips[:8].tolist()], '| units', sorted(set(flips[:, 1].tolist()))[:16])
pre32 = (x @ w.t() + b)
print('torch fp32 (cuBLAS) flips vs float64:', int(((pre32 > 0) != (pre64 > 0)).sum()))
