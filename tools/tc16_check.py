"""GPU probe of the tensor-core engines (3xTF32 `tc32`, 3xFP16 `tc16`) against float64 / the SIMT path.
Run under `timeout` on the GPU box; every stage prints as it finishes so a hang is localised."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops, synth  # noqa: E402

torch.cuda.set_device(0)
MODES = [m for m in os.environ.get('SGG_CHECK_MODES', 'simt,tc32,tc16').split(',') if m]


T0 = time.time()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3     # us


def linear_case(M, N, K, relu=False, seed=0, scale=1.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(M, K, device='cuda', generator=g) * scale
    w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    ref = x.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    msg = 'linear M=%5d N=%5d K=%5d scale=%g |' % (M, N, K, scale)
    for mode in MODES:
        ops.set_gemm_mode(mode)
        y = ops.linear(x, w, b, relu=relu); torch.cuda.synchronize()
        err = (y.double() - ref).abs().max().item()
        us = timeit(lambda: ops.linear(x, w, b, relu=relu))
        msg += ' %s err %.2e %.1f us (%.0f TF)' % (mode, err, us, 2.0 * M * N * K / us / 1e6)
    print(msg, flush=True)


def mp_case(B, boxes, edges, T, seed):
    g = synth.synth_graph(B, boxes, edges, seed)
    N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
    of, ef = synth.synth_l1_feats(N, E, seed)
    p = {k: torch.from_numpy(v).cuda() for k, v in synth.synth_params(111, level='l1').items()}
    rel = torch.from_numpy(np.ascontiguousarray(g['rel_inds'][:, 1:3])).cuda()
    gr = ops.build_graph(rel, N, validate=True)
    o, e = torch.from_numpy(of).cuda(), torch.from_numpy(ef).cuda()
    outs = {}
    msg = 'l1 B=%d N=%d E=%d T=%d |' % (B, N, E, T)
    for mode in MODES:
        ops.set_gemm_mode(mode)
        obj_rep = ops.linear(o, p['obj_unary.weight'], p['obj_unary.bias'])
        rel_rep = ops.linear(e, p['edge_unary.weight'], p['edge_unary.bias'], relu=True)
        V, Eh = ops.message_pass(rel_rep, obj_rep, gr, p, T)
        od, rd = ops.l1_forward(o, e, gr, p, T)
        torch.cuda.synchronize()
        outs[mode] = (V.clone(), Eh.clone(), od.clone(), rd.clone())
        P_t = ops.linear(obj_rep, p['edge_gru.weight_ih'])
        gates = torch.rand(E, 4, device='cuda')
        eg = ops.edge_gru(rel_rep, P_t, gates, gr, p); torch.cuda.synchronize()
        outs[mode] += (eg.clone(),)
        t_mp = timeit(lambda: ops.message_pass(rel_rep, obj_rep, gr, p, T), 10)
        t_eg = timeit(lambda: ops.edge_gru(rel_rep, P_t, gates, gr, p), 20)
        plan = ops.L1Plan(p, N, E, 4096, T, 'cuda')
        t_l1 = timeit(lambda: plan.run(o, e, gr), 10)
        msg += ' %s: mp %.0f us, edge_gru %.1f us, l1 %.0f us;' % (mode, t_mp, t_eg, t_l1)
    print(msg, '[%.1fs]' % (time.time() - T0), flush=True)
    base = MODES[0]
    for mode in MODES[1:]:
        d = [float((a - b).abs().max()) for a, b in zip(outs[mode], outs[base])]
        print('   %s vs %s: max|d| V %.2e Eh %.2e obj_dists %.2e rel_dists %.2e edge_gru %.2e'
              % (mode, base, d[0], d[1], d[2], d[3], d[4]), flush=True)


if __name__ == '__main__':
    print('modes', MODES, 'lib default engine', ops.tc_engine(), flush=True)
    for shape in [] if os.environ.get('SGG_CHECK_SKIP_LINEAR') else [(128, 64, 64), (128, 64, 512), (100, 151, 512), (240, 1536, 512), (2400, 51, 512), (240, 512, 4096),
                  (2400, 512, 4096), (9600, 512, 4096), (1000, 4096, 4096)]:
        linear_case(*shape)
    if not os.environ.get('SGG_CHECK_SKIP_LINEAR'):
        linear_case(300, 512, 4096, relu=True)
        linear_case(300, 512, 4096, scale=100.0)
        linear_case(300, 512, 4096, scale=1e-4)
    print('linear cases done at %.1fs' % (time.time() - T0), flush=True)
    mp_case(1, 10, 90, 3, 7)
    mp_case(8, 30, 300, 3, 1236)
    mp_case(32, 30, 300, 3, 1237)
    mp_case(8, 64, 2000, 6, 1238)
    print('tc16 check done', flush=True)
