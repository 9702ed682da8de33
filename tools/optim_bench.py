"""Training-tail sweeps at the full model size (247.75 M trainable fp32 parameters, the shapes of SURVEY.md §8a):
gradient-norm sweep, fused clip+SGD sweep (csrc/train.cu) against the HBM roofline, next to what the reference runs on a
GPU (its per-tensor clip_grad_norm loop + torch.optim.SGD).  Prints one JSON line."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import optim, _lib
from sgg_b200.ops import _ptr, _stream

torch.cuda.set_device(0)
shapes = {}
for pre in ('roi_fmap.1.', 'roi_fmap_obj.'):
    shapes[pre + '0.weight'] = (4096, 25088); shapes[pre + '0.bias'] = (4096,)
    shapes[pre + '3.weight'] = (4096, 4096); shapes[pre + '3.bias'] = (4096,)
shapes.update({'union_boxes.conv.0.weight': (256, 2, 7, 7), 'union_boxes.conv.0.bias': (256,),
               'union_boxes.conv.2.weight': (256,), 'union_boxes.conv.2.bias': (256,),
               'union_boxes.conv.4.weight': (512, 256, 3, 3), 'union_boxes.conv.4.bias': (512,),
               'union_boxes.conv.6.weight': (512,), 'union_boxes.conv.6.bias': (512,),
               'rel_fc.weight': (51, 512), 'rel_fc.bias': (51,), 'obj_fc.weight': (151, 512), 'obj_fc.bias': (151,),
               'obj_unary.weight': (512, 4096), 'obj_unary.bias': (512,), 'edge_unary.weight': (512, 4096), 'edge_unary.bias': (512,)})
for g in ('edge_gru', 'node_gru'):
    shapes[g + '.weight_ih'] = (1536, 512); shapes[g + '.weight_hh'] = (1536, 512)
    shapes[g + '.bias_ih'] = (1536,); shapes[g + '.bias_hh'] = (1536,)
for k in ('sub_vert', 'obj_vert', 'out_edge', 'in_edge'):
    shapes[k + '_w_fc.0.weight'] = (1, 1024); shapes[k + '_w_fc.0.bias'] = (1,)
gen = torch.Generator(device='cuda').manual_seed(0)
named = [(n, torch.nn.Parameter(0.02 * torch.randn(s, device='cuda', generator=gen))) for n, s in shapes.items()]
n_params = sum(p.numel() for _, p in named)
for _, p in named:
    p.grad = 1e-3 * torch.randn(p.shape, device='cuda', generator=gen)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']


def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


fc = [p for n, p in named if n.startswith('roi_fmap')]; rest = [p for n, p in named if not n.startswith('roi_fmap')]
opt = optim.FusedSGD([{'params': fc, 'lr': 1e-4}, {'params': rest}], lr=1e-3, momentum=0.9, weight_decay=1e-4,
                     emit_operand_split=False)
res = {'params': n_params, 'tensors': len(named), 'hbm_peak_gbs': peak}
lib = _lib.load()
n0 = lib.sgg_launch_count()
opt.step(max_norm=5.0)
res['launches_per_step'] = int(lib.sgg_launch_count() - n0)
t_step = timeit(lambda: opt.step(max_norm=5.0))
tab = opt._table
t_norm = timeit(lambda: lib.sgg_mt_grad_norm(_ptr(tab.dev), tab.n, tab.chunks, 5.0, _ptr(tab.norm), _ptr(tab.ws), tab.ws.numel(), _stream()))
t_sgd = timeit(lambda: lib.sgg_mt_sgd_step(_ptr(tab.dev), tab.n, tab.chunks, _ptr(tab.norm), 0.9, 0, _stream()))
res['fused'] = {'step_ms': t_step, 'norm_ms': t_norm, 'sgd_ms': t_sgd,
                'norm_gbs': 4.0 * n_params / t_norm / 1e6, 'sgd_gbs': 20.0 * n_params / t_sgd / 1e6,
                'step_gbs': 24.0 * n_params / t_step / 1e6}
res['fused']['sgd_frac_of_hbm_peak'] = res['fused']['sgd_gbs'] / peak
res['fused']['norm_frac_of_hbm_peak'] = res['fused']['norm_gbs'] / peak
res['fused']['step_frac_of_hbm_peak'] = res['fused']['step_gbs'] / peak

# what the reference does on a GPU: python loop of per-tensor norms + mul_, then torch.optim.SGD (foreach)
ref = torch.optim.SGD([{'params': fc, 'lr': 1e-4}, {'params': rest}], lr=1e-3, momentum=0.9, weight_decay=1e-4)
def ref_step():
    total = 0
    for _, p in named:
        total += p.grad.data.norm(2) ** 2
    total = total ** 0.5
    coef = 5.0 / (total + 1e-6)
    if coef < 1:
        for _, p in named:
            p.grad.data.mul_(coef)
    ref.step()
if not os.environ.get('OPTIM_BENCH_SKIP_REF'):
    res['torch_reference_style'] = {'step_ms': timeit(ref_step, reps=5, warm=2)}
    res['speedup_vs_torch'] = res['torch_reference_style']['step_ms'] / t_step
print(json.dumps(res))
