"""fc6 of the edge branch (9600 x 25088 -> 4096) on pre-split operands, in a loop: ncu target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops
torch.cuda.set_device(0)
g = torch.Generator(device='cuda').manual_seed(0)
M, K, N = int(os.environ.get('FM', 9600)), 25088, 4096
x = torch.rand(M, K, device='cuda', generator=g)
w = torch.nn.Parameter(torch.randn(N, K, device='cuda', generator=g) / 158.0)
b = torch.zeros(N, device='cuda')
hi = x.half(); pl = torch.stack((hi, ((x - hi.float()) * 2048).half())).contiguous()
for _ in range(3):
    y = ops.linear(None, w, b, relu=True, x_planes=pl)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    y = ops.linear(None, w, b, relu=True, x_planes=pl)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 5
print('fc6 pre-split: %.3f ms, %.0f fp32-equiv TFLOP/s' % (ms, 2.0 * M * K * N / ms / 1e9))
