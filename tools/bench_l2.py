"""Stage timings (CUDA events) of the feature-head path — SURVEY.md §8f rank 1 / BASELINE cfg3 shapes:
RoIAlign (objects + union boxes), union-box geometry, fc6/fc7 heads, then the L1 path.  Prints a JSON line."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops, synth

B = int(os.environ.get('B', 32)); NB = 30; NE = 300
torch.cuda.set_device(0)
g = synth.synth_graph(B, NB, NE, 77)
N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
fmap = torch.relu(torch.randn(B, 512, 38, 38, device='cuda'))
rois = dev(g['rois']); rel = dev(g['rel_inds'])
gen = torch.Generator(device='cuda').manual_seed(0)
P = {}
def u(*shape, fan):
    return (torch.rand(*shape, device='cuda', generator=gen) * 2 - 1) / fan ** 0.5
for pre in ('roi_fmap.1.', 'roi_fmap_obj.'):
    P[pre + '0.weight'] = u(4096, 25088, fan=25088); P[pre + '0.bias'] = u(4096, fan=25088)
    P[pre + '3.weight'] = u(4096, 4096, fan=4096); P[pre + '3.bias'] = u(4096, fan=4096)
l1 = {k: dev(v) for k, v in synth.synth_params(5, level='l2').items() if not k.startswith('roi_fmap')}
P.update(l1)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out

res = {'B': B, 'N': N, 'E': E}
for mode in ('warm', 'tc', 'simt'):       # 'warm': one untimed pass so the allocator / weight-split caches are in steady state
    ops.set_gemm_mode('tc' if mode == 'warm' else mode)
    r = {}
    r['roi_align_ms'], (nf, ef) = timeit(lambda: ops.node_edge_features(fmap, rois, rel[:, 1:3]))
    r['union_geom_add_ms'], ef2 = timeit(lambda: ops.union_geom(rois, rel[:, 1:3], P, ef))
    # eval forward: geometry embedding [E,C] first, added inside the RoIAlign kernel (one write of [E,C,7,7])
    r['fused_geom_roi_align_ms'], (nf, ef2) = timeit(
        lambda: ops.node_edge_features(fmap, rois, rel[:, 1:3], edge_add=ops.union_geom(rois, rel[:, 1:3], P)))
    r['fc6_edge_ms'], h = timeit(lambda: ops.linear(ef2.view(E, -1), P['roi_fmap.1.0.weight'], P['roi_fmap.1.0.bias'], relu=True), 3)
    r['fc7_edge_ms'], e4096 = timeit(lambda: ops.linear(h, P['roi_fmap.1.3.weight'], P['roi_fmap.1.3.bias']))
    r['fc6_node_ms'], hn = timeit(lambda: ops.linear(nf.view(N, -1), P['roi_fmap_obj.0.weight'], P['roi_fmap_obj.0.bias'], relu=True), 3)
    r['fc7_node_ms'], n4096 = timeit(lambda: ops.linear(hn, P['roi_fmap_obj.3.weight'], P['roi_fmap_obj.3.bias'], relu=True))
    gr = ops.build_graph(rel[:, 1:3], N)
    r['l1_ms'], _ = timeit(lambda: ops.l1_forward(n4096, e4096, gr, P, 3))
    r['fc6_edge_tflops'] = 2.0 * E * 25088 * 4096 / (r['fc6_edge_ms'] * 1e-3) / 1e12
    r['roi_align_gbs'] = (N + E) * 512 * 49 * 4 / (r['roi_align_ms'] * 1e-3) / 1e9
    r['total_ms'] = sum(v for k, v in r.items() if k.endswith('_ms') and k not in ('roi_align_ms', 'union_geom_add_ms'))
    r['total_unfused_ms'] = sum(v for k, v in r.items() if k.endswith('_ms') and k not in ('fused_geom_roi_align_ms', 'total_ms'))
    r['images_per_s'] = B / (r['total_ms'] * 1e-3)
    if mode != 'warm':
        res[mode] = r
ops.set_gemm_mode('tc')
print(json.dumps(res))
