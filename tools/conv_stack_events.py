"""VGG16 conv stack (ops.vgg_features order) with a CUDA event between consecutive launches of ONE pass: per-layer ms inside the stack."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops, _lib
from sgg_b200.ops import _ptr, _stream, check
from sgg_b200.model import _vgg16_parts
torch.cuda.set_device(0)
torch.manual_seed(0)
B, H, W = int(os.environ.get('CB', 32)), 608, 608
feats, _ = _vgg16_parts()
layers = ops.vgg_layers(feats.cuda().eval())
lib = _lib.load()
img = torch.rand(B, 3, H, W, device='cuda')
wps = [None] + [ops.conv_weight_planes(wt) for wt, _, _ in layers[1:]]
# preallocate every activation buffer
bufs = [torch.empty((2, B, H, W, 64), dtype=torch.float16, device='cuda')]
h, w = H, W
for wt, bs, pool in layers[1:]:
    if pool: h //= 2; w //= 2
    bufs.append(torch.empty((2, B, h, w, wt.shape[0]), dtype=torch.float16, device='cuda'))
def one_pass(record):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(layers) + 1)]
    w0, b0, _ = layers[0]
    if record: ev[0].record()
    check(lib.sgg_conv3x3_first(_ptr(img), _ptr(w0), _ptr(b0), B, H, W, 64, _ptr(bufs[0]), _stream()), 'first')
    if record: ev[1].record()
    h, w, cin = H, W, 64
    for li, (wt, bs, pool) in enumerate(layers[1:], start=1):
        cout = wt.shape[0]
        check(lib.sgg_conv3x3_tc(_ptr(bufs[li - 1]), _ptr(wps[li]), _ptr(bs), B, h, w, cin, cout, 1, 1 if pool else 0, _ptr(bufs[li]), None, _stream()), 'conv')
        if record: ev[li + 1].record()
        if pool: h //= 2; w //= 2
        cin = cout
    return ev
one_pass(False); torch.cuda.synchronize()
for rep in range(2):
    ev = one_pass(True); torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(layers))]
    print('pass %d: first %.3f | ' % (rep, ts[0]) + ' '.join('%.3f' % t for t in ts[1:]) + ' | total %.2f ms' % sum(ts))
