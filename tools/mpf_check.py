"""GPU validation + timing of the fused message-passing path (csrc/mp_fused.cu) against the SIMT CUDA path and the numpy
oracle.  Env toggles are per process: SGG_MP_FUSED=0 (old 7-launch schedule), SGG_MPF_BK=32, SGG_MPF_PDL=1.
Every stage prints as it finishes, so a trap / hang is localised.  Run under `timeout`."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops, synth, _lib  # noqa: E402

torch.cuda.set_device(0)
lib = _lib.load()
TAG = 'fused=%s bk=%s pdl=%s' % (os.environ.get('SGG_MP_FUSED', '1'), os.environ.get('SGG_MPF_BK', '32'), os.environ.get('SGG_MPF_PDL', '1'))


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def case(B, boxes, edges, T, seed, scale=1.0, ragged=False, tape=True, oracle=False):
    g = synth.synth_graph(B, boxes, edges, seed, ragged=ragged) if ragged else synth.synth_graph(B, boxes, edges, seed)
    N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
    of, ef = synth.synth_l1_feats(N, E, seed)
    pn = synth.synth_params(111, level='l1')
    if scale != 1.0:
        pn = {k: (v * scale if v.ndim == 2 and 'unary' not in k else v) for k, v in pn.items()}
    p = {k: torch.from_numpy(v).cuda() for k, v in pn.items()}
    rel = torch.from_numpy(np.ascontiguousarray(g['rel_inds'][:, 1:3])).cuda()
    gr = ops.build_graph(rel, N, validate=True)
    o, e = torch.from_numpy(of).cuda(), torch.from_numpy(ef).cuda()
    res = {}
    for mode in ('simt', 'tc16'):
        ops.set_gemm_mode(mode)
        obj_rep = ops.linear(o, p['obj_unary.weight'], p['obj_unary.bias'])
        rel_rep = ops.linear(e, p['edge_unary.weight'], p['edge_unary.bias'], relu=True)
        V, Eh = ops.message_pass(rel_rep, obj_rep, gr, p, T)
        od, rd = ops.l1_forward(o, e, gr, p, T)
        torch.cuda.synchronize()
        r = dict(V=V, Eh=Eh, od=od.clone(), rd=rd.clone())
        if tape:
            Vt, Et, tp = ops.message_pass_train(rel_rep, obj_rep, gr, p, T)
            torch.cuda.synchronize()
            r.update(Vt=Vt, Et=Et, tape=tp)
        res[mode] = r
    a, b = res['simt'], res['tc16']
    msg = '[%s] B=%d N=%d E=%d T=%d scale=%g |' % (TAG, B, N, E, T, scale)
    for k in ('V', 'Eh', 'od', 'rd') + (('Vt', 'Et', 'tape') if tape else ()):
        d = (a[k].double() - b[k].double()).abs().max().item() if a[k].numel() else 0.0
        msg += ' %s %.2e' % (k, d)
    if tape:      # per-section tape diffs (states | cacheV | cacheE | gates | ctx | P)
        H = 512
        al = lambda n: (n + 63) // 64 * 64
        secs = [('states', (T + 1) * (N + E) * H), ('cacheV', (T + 1) * N * 4 * H), ('cacheE', (T + 1) * E * 4 * H),
                ('gates', T * E * 4), ('ctx', T * N * H), ('P', T * N * 3 * H)]
        off = 0
        for nm, n in secs:
            d = (a['tape'][off:off + n].double() - b['tape'][off:off + n].double()).abs().max().item() if n else 0.0
            msg += ' %s %.1e' % (nm, d)
            off += al(n)
    if oracle:
        from oracle import imp_numpy as O
        od_ref, rd_ref = O.l1_forward(of, ef, g['rel_inds'][:, 1:3], pn, T)
        msg += ' | vs oracle od %.2e rd %.2e' % (np.abs(b['od'].cpu().numpy() - od_ref).max(), np.abs(b['rd'].cpu().numpy() - rd_ref).max())
    print(msg, flush=True)
    return p, gr, o, e, N, E


def phases(tag, n_ctas, which=0, sel=None):
    buf = (C.c_longlong * (8 * n_ctas))()
    assert lib.sgg_mpf_debug_timing(buf, n_ctas, which) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(n_ctas, 8)
    if sel is not None:
        t = t[sel]; n_ctas = t.shape[0]
    names = ['setup', 'first data', 'main loop', 'phase1 TMEM->smem', 'barrier', 'phase2 pointwise', 'tail']
    d = np.stack([t[:, i + 1] - t[:, i] for i in range(7)], 1)
    tot = t[:, 7] - t[:, 0]
    print('%s: %d CTAs total mean %.0f max %.0f | ' % (tag, n_ctas, tot.mean(), tot.max()) +
          ' '.join('%s %.0f' % (nm, d[:, k].mean()) for k, nm in enumerate(names)), flush=True)


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    print('start', TAG, flush=True)
    if what in ('all', 'check'):
        case(1, 10, 90, 3, 7, oracle=True)
        case(2, 30, 300, 3, 11, oracle=True)
        case(8, 30, 300, 3, 1236)
        case(8, 30, 300, 3, 1236, scale=3.0)
        case(8, 64, 2000, 6, 99, tape=False)
        case(32, 30, 300, 3, 5, tape=False)
        try:
            case(6, 20, 150, 6, 3, ragged=True)
        except TypeError:
            pass
    if what in ('all', 'time'):
        p, gr, o, e, N, E = case(8, 30, 300, 3, 1236, tape=False)
        ops.set_gemm_mode('tc16')
        obj_rep = ops.linear(o, p['obj_unary.weight'], p['obj_unary.bias'])
        rel_rep = ops.linear(e, p['edge_unary.weight'], p['edge_unary.bias'], relu=True)
        print('[%s] message_pass T=3 cfg2: %.1f us' % (TAG, timeit(lambda: ops.message_pass(rel_rep, obj_rep, gr, p, 3))), flush=True)
        print('[%s] message_pass T=1 cfg2: %.1f us' % (TAG, timeit(lambda: ops.message_pass(rel_rep, obj_rep, gr, p, 1))), flush=True)
        print('[%s] message_pass T=0 cfg2: %.1f us' % (TAG, timeit(lambda: ops.message_pass(rel_rep, obj_rep, gr, p, 0))), flush=True)
        plan = ops.L1Plan(p, N, E, 4096, 3, 'cuda')
        rel = None
        print('[%s] l1 eager cfg2: %.1f us' % (TAG, timeit(lambda: plan.run(o, e, gr))), flush=True)
        gph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.graph(gph, stream=s):
            plan.run(o, e, gr)
        print('[%s] l1 cuda-graph cfg2: %.1f us' % (TAG, timeit(lambda: gph.replay())), flush=True)
        if os.environ.get('SGG_TC_TIMING') and os.environ.get('SGG_MP_FUSED', '1') != '0':
            ops.message_pass(rel_rep, obj_rep, gr, p, 1); torch.cuda.synchronize()
            phases('k_mp_gru EDGE tiles (launch B, last iteration: no planes / partials)', 147, sel=slice(0, 133))
            phases('k_mp_gru NODE tiles (launch B, last iteration)', 147, sel=slice(133, 147))
            phases('k_mp_pre LIN tiles (launch A)', 148, which=1, sel=slice(0, 48))
            phases('k_mp_pre CTX CTAs (launch A)', 148, which=1, sel=slice(48, 148))
            ops.message_pass(rel_rep, obj_rep, gr, p, 2); torch.cuda.synchronize()
            # after T = 2 the buffer still holds launch B of the LAST iteration; the first iteration's B emitted planes +
            # partials, as the INIT launch does: report INIT (T = 0) for that variant
            ops.message_pass(rel_rep, obj_rep, gr, p, 0); torch.cuda.synchronize()
            phases('k_mp_gru INIT edge tiles (planes + partials emitted)', 147, sel=slice(0, 133))
    print('done', TAG, flush=True)
