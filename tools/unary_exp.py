"""Experiment: edge-unary GEMM (2400 x 4096 -> 512) — fp32-input LINEAR vs split pass + pre-split kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops
torch.cuda.set_device(0)
g = torch.Generator(device='cuda').manual_seed(0)
M, K, N = 2400, 4096, 512
xs = [torch.randn(M, K, device='cuda', generator=g) for _ in range(6)]      # 6 x 39 MB > L2 with the planes
w = torch.nn.Parameter(torch.randn(N, K, device='cuda', generator=g) / 64.0)
b = torch.zeros(N, device='cuda')
pls = []
for x in xs:
    hi = x.half(); pls.append(torch.stack((hi, ((x - hi.float()) * 2048).half())).contiguous())
def timeit(fn, reps=60):
    for i in range(6): fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(reps): fn(i)
    gr.replay(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) * 1e3 / reps
t32 = timeit(lambda i: ops.linear(xs[i % 6], w, b, relu=True))
tpre = timeit(lambda i: ops.linear(None, w, b, relu=True, x_planes=pls[i % 6], out_planes=True))
ys = [torch.empty_like(x) for x in xs]
tcopy = timeit(lambda i: torch.add(xs[i % 6], 1.0, out=ys[i % 6]))
print('edge-unary 2400x4096->512: fp32-input LINEAR %.1f us | pre-split kernel %.1f us | 39 MB read + 39 MB write pass %.1f us' % (t32, tpre, tcopy))
