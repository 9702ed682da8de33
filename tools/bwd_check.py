"""GPU probe: tensor-core backward building blocks (3xTF32 GEMM on transposed operands) against float64."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops
torch.cuda.set_device(0)
g = torch.Generator(device='cuda').manual_seed(1)
print('--- tc32 linear y = x w^T: (M, N, K)')
for (M, N, K) in [(512, 4096, 2400), (512, 4096, 3008), (512, 4096, 4800), (512, 4096, 9600), (512, 4096, 9728),
                  (512, 4096, 8192), (128, 128, 9600), (512, 512, 9600), (1536, 512, 9600), (4096, 25088, 1024)]:
    x = torch.randn(M, K, device='cuda', generator=g); w = torch.randn(N, K, device='cuda', generator=g)
    wT = ops._transpose(w.t().contiguous(), K, N, N, K, split=True)     # split(w) through the transpose kernel: [2, N, K]
    y = ops._tc32_linear(x, wT, M, N, K)
    ref = x.double() @ w.double().t()
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    print('M=%5d N=%5d K=%5d  rel err %.2e' % (M, N, K, err), flush=True)
print('--- linear_backward (M, Nout, K)')
for (M, N, K) in [(2400, 512, 4096), (9600, 512, 4096), (9600, 51, 512), (9600, 4096, 4096), (960, 4096, 4096)]:
    x = torch.randn(M, K, device='cuda', generator=g); w = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    dy = torch.randn(M, N, device='cuda', generator=g)
    dx, dw, db = ops.linear_backward(x, w, dy)
    rx, rw, rb = dy.double() @ w.double(), dy.double().t() @ x.double(), dy.double().sum(0)
    print('M=%5d N=%5d K=%5d  dx %.2e dw %.2e db %.2e' % (M, N, K, (dx.double() - rx).abs().max().item() / rx.abs().max().item(),
          (dw.double() - rw).abs().max().item() / rw.abs().max().item(), (db.double() - rb).abs().max().item() / rb.abs().max().item()), flush=True)
