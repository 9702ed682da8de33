"""Probe of the tcgen05 3xTF32 GEMM engine against float64 (run under `timeout` on the GPU box)."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops

torch.cuda.set_device(0)
def run(M, N, K, relu=False, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(M, K, device='cuda', generator=g); w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    b = torch.randn(N, device='cuda', generator=g)
    ref = x.double() @ w.double().t() + b.double()
    if relu: ref = ref.clamp_min(0)
    out = {}
    for mode in ('simt', 'tc'):
        ops.set_gemm_mode(mode)
        y = ops.linear(x, w, b, relu=relu); torch.cuda.synchronize()
        out[mode] = (y.double() - ref).abs().max().item()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10): ops.linear(x, w, b, relu=relu)
        t1.record(); torch.cuda.synchronize()
        out[mode + '_us'] = t0.elapsed_time(t1) * 100
    torch.backends.cuda.matmul.allow_tf32 = False
    yt = torch.addmm(b, x, w.t())
    out['torch_fp32'] = (yt.double() - ref).abs().max().item()
    print('M=%5d N=%5d K=%5d  err simt %.2e tc %.2e torch %.2e | simt %.1f us tc %.1f us  (%.1f TFLOP/s tc)'
          % (M, N, K, out['simt'], out['tc'], out['torch_fp32'], out['simt_us'], out['tc_us'],
             2.0 * M * N * K / out['tc_us'] / 1e6), flush=True)
    return out

for shape in [(128, 64, 32), (128, 64, 512), (100, 151, 512), (2400, 512, 4096), (240, 1536, 512), (2400, 51, 512),
              (9600, 512, 4096), (1000, 4096, 4096)]:
    run(*shape)
run(300, 512, 4096, relu=True)
print('tc probe done')
