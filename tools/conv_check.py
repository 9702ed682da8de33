"""GPU probe of the tcgen05 VGG16 conv stack (csrc/conv_tc.cu) against torchvision / cuDNN fp32 (allow_tf32=False)."""
import os, sys, time
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops
from sgg_b200.model import _vgg16_parts
torch.cuda.set_device(0)
torch.manual_seed(0)
feats, _ = _vgg16_parts()
feats = feats.cuda().eval()
layers = ops.vgg_layers(feats)
print('layers:', [(tuple(w.shape), p) for w, _, p in layers], flush=True)


def ref(x, mods):
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        return mods(x)


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# correctness on prefixes of the stack (localises a failing layer), odd sizes cover partial tiles
mods = list(feats.children())
for (H, W) in [(32, 48), (96, 160)]:
    x = torch.rand(2, 3, H, W, device='cuda')
    n_conv = 0
    for i, m in enumerate(mods):
        if not isinstance(m, nn.Conv2d):
            continue
        n_conv += 1
        # prefix = up to and including this conv's ReLU and a following pool
        j = i + 2
        if j < len(mods) and isinstance(mods[j], nn.MaxPool2d):
            j += 1
        prefix = nn.Sequential(*mods[:j])
        lay = layers[:n_conv]
        if n_conv == 1:
            continue                      # the first layer alone emits planes only; checked through the 2-layer prefix
        r = ref(x, prefix)
        try:
            y = ops.vgg_features(x, lay)
            torch.cuda.synchronize()
            err = (y - r).abs().max().item() / max(1e-6, r.abs().max().item())
            print('HxW %dx%d  convs=%2d  out %s  rel err %.2e  (max |ref| %.3f)' % (H, W, n_conv, tuple(y.shape), err, r.abs().max().item()), flush=True)
        except Exception as ex:
            print('HxW %dx%d convs=%d FAILED: %s' % (H, W, n_conv, str(ex)[:200]), flush=True)
            break
print('overflow flag:', ops._lib.load().sgg_conv_overflow(1))
for B in (4, 32):
    x = torch.rand(B, 3, 608, 608, device='cuda')
    t_ref = timeit(lambda: ref(x, feats), 2)
    t_tc = timeit(lambda: ops.vgg_features(x, layers), 2)
    y = ops.vgg_features(x, layers); r = ref(x, feats)
    err = (y - r).abs().max().item() / r.abs().max().item()
    gf = 2 * 15.35e9 * (608 / 224) ** 2 * B / 1e9
    print('B=%d 608x608: cuDNN fp32 %.2f ms (%.0f TF) | tcgen05 3xFP16 %.2f ms (%.0f fp32-equiv TF) | rel err %.2e | %.0f img/s vs %.0f img/s'
          % (B, t_ref, gf / t_ref, t_tc, gf / t_tc, err, B / t_tc * 1e3, B / t_ref * 1e3), flush=True)
