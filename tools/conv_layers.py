"""Per-layer timing of the tcgen05 VGG16 conv stack (csrc/conv_tc.cu): ms, fp32-equivalent TFLOP/s and cycles per k-block."""
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
h, w, cin = H, W, 64
tot = 0.0
for li, (wt, bs, pool) in enumerate(layers[1:], start=1):
    cout = wt.shape[0]
    x = (torch.rand((2, B, h, w, cin), device='cuda') * 0.1).half()
    wp = ops.conv_weight_planes(wt)
    ho, wo = (h // 2, w // 2) if pool else (h, w)
    out = torch.empty((2, B, ho, wo, cout), dtype=torch.float16, device='cuda')
    def run():
        check(lib.sgg_conv3x3_tc(_ptr(x), _ptr(wp), _ptr(bs), B, h, w, cin, cout, 1, 1 if pool else 0, _ptr(out), None, _stream()), 'conv')
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    if os.environ.get('CEACH'):
        each = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); run(); e1.record(); torch.cuda.synchronize(); each.append(round(e0.elapsed_time(e1), 3))
        print('   single launches (ms):', each, ' free/total GB: %.1f / %.1f' % tuple(v / 1e9 for v in torch.cuda.mem_get_info()))
    fl = 2.0 * B * h * w * cout * 9 * cin
    tiles = ((h + 7) // 8) * ((w + 15) // 16) * B * (cout // (128 if cout % 128 == 0 else 64))
    kb = 9 * cin // 64
    waves = tiles / 148.0
    print('L%02d %3d->%3d %3dx%3d pool=%d  %6.3f ms  %5.0f TF  tiles %6d x %3d kb, %.0f ns per tile-kblock-wave' % (
        li, cin, cout, h, w, int(pool), ms, fl / ms / 1e9, tiles, kb, ms * 1e6 / (waves * kb)), flush=True)
    tot += ms
    h, w, cin = ho, wo, cout
    del x, out
print('total %.2f ms' % tot)
