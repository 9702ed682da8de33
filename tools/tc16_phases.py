"""Per-CTA phase breakdown (clock64) of the 3xFP16 kernels.  Run with SGG_TC_TIMING=1 on the GPU box."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('SGG_TC_TIMING', '1')
from sgg_b200 import ops, synth, _lib
lib = _lib.load()
ops.set_gemm_mode('tc16')
NAMES = ['setup', 'first data', 'main loop (to tmem_full)', 'phase 1 (TMEM->smem)', 'barrier', 'phase 2 (pointwise)', 'tail']

def report(tag, n_ctas, gru=True, skip0=False):
    buf = (C.c_longlong * (8 * n_ctas))()
    assert lib.sgg_tc_debug_timing(buf, n_ctas) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(n_ctas, 8)
    if skip0:
        t = t[1:]; n_ctas -= 1
    if gru:
        d = np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4],
                      t[:, 6] - t[:, 5], t[:, 7] - t[:, 6]], 1)
        tot = t[:, 7] - t[:, 0]
        print('%s: %d CTAs, total cycles mean %.0f max %.0f' % (tag, n_ctas, tot.mean(), tot.max()))
        for k, nm in enumerate(NAMES):
            print('   %-28s mean %8.0f  min %8.0f  max %8.0f' % (nm, d[:, k].mean(), d[:, k].min(), d[:, k].max()))
    else:
        tot = t[:, 7] - t[:, 0]
        print('%s: %d CTAs, total cycles mean %.0f max %.0f; setup %.0f; first data %.0f'
              % (tag, n_ctas, tot.mean(), tot.max(), (t[:, 1] - t[:, 0]).mean(), (t[:, 2] - t[:, 1]).mean()))
    sys.stdout.flush()

g = synth.synth_graph(8, 30, 300, 1236)
N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
of, ef = synth.synth_l1_feats(N, E, 1236)
p = {k: torch.from_numpy(v).cuda() for k, v in synth.synth_params(111, level='l1').items()}
rel = torch.from_numpy(np.ascontiguousarray(g['rel_inds'][:, 1:3])).cuda()
gr = ops.build_graph(rel, N)
o, e = torch.from_numpy(of).cuda(), torch.from_numpy(ef).cuda()
obj_rep = ops.linear(o, p['obj_unary.weight'], p['obj_unary.bias'])
for _ in range(3):
    rel_rep = ops.linear(e, p['edge_unary.weight'], p['edge_unary.bias'], relu=True)
report('edge unary LINEAR 2400x512x4096 (stream-K, 148 CTAs)', 148, gru=False)
P_t = ops.linear(obj_rep, p['edge_gru.weight_ih'])
report('P LINEAR 240x1536x512', 48, gru=False)
gates = torch.rand(E, 4, device='cuda')
for _ in range(3):
    ops.edge_gru(rel_rep, P_t, gates, gr, p)
report('EDGE GRU E=2400 (grid 7x19)', 133)
# node GRU in (near) isolation: a 1-edge graph keeps the concurrent edge-branch kernels to a single CTA (CTA 0 is skipped)
gr1 = ops.build_graph(rel[:1].contiguous(), N)
V, Eh = ops.message_pass(rel_rep[:1].contiguous(), obj_rep, gr1, p, 1)
report('NODE GRU N=240 (grid 16x2) [message_pass T=1 on a 1-edge graph]', 32, skip0=True)
