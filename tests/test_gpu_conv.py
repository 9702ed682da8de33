"""GPU: the tcgen05 VGG16 conv stack (csrc/conv_tc.cu; rel_model_base.py:184, :310-321) against torchvision's own modules
run through cuDNN in fp32 (allow_tf32=False) — the library path the reference uses.  Bar: 1e-4 relative to the output scale
(north-star tolerance), measured ~1e-5 after 13 layers."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _features():
    from sgg_b200.model import _vgg16_parts
    torch.manual_seed(0)
    feats, _ = _vgg16_parts()
    return feats.cuda().eval()


def _ref(x, mods):
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        return mods(x)


def test_vgg_layers_recognises_the_reference_backbone():
    from sgg_b200 import ops
    lay = ops.vgg_layers(_features())
    assert lay is not None and len(lay) == 13
    assert [p for _, _, p in lay] == [False, True, False, True, False, False, True, False, False, True, False, False, False]
    assert ops.vgg_layers(nn.Sequential(nn.Conv2d(3, 64, 5, padding=2), nn.ReLU())) is None      # not 3x3: library path


@pytest.mark.parametrize('shape', [(2, 32, 48), (1, 96, 160), (3, 48, 16)])
def test_conv_stack_prefixes_vs_cudnn_fp32(shape):
    """Every prefix of the stack (so each layer type is compared where it is the last one: planes -> fp32 NCHW, with and
    without the fused max-pool, Cout = 64 / 128 / 256 / 512), on sizes with partial 8 x 16 tiles."""
    from sgg_b200 import ops
    feats = _features()
    layers = ops.vgg_layers(feats)
    mods = list(feats.children())
    B, H, W = shape
    x = torch.rand(B, 3, H, W, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))
    n_conv = 0
    for i, m in enumerate(mods):
        if not isinstance(m, nn.Conv2d):
            continue
        n_conv += 1
        if n_conv == 1:
            continue
        j = i + 2
        if j < len(mods) and isinstance(mods[j], nn.MaxPool2d):
            j += 1
        r = _ref(x, nn.Sequential(*mods[:j]))
        y = ops.vgg_features(x, layers[:n_conv])
        assert y.shape == r.shape
        err = float((y - r).abs().max()) / max(1e-6, float(r.abs().max()))
        assert err <= 1e-4, 'prefix of %d convs: rel err %.3e' % (n_conv, err)
    assert ops._lib.load().sgg_conv_overflow(1) == 0


def test_conv_stack_large_activations_raise_the_range_flag():
    from sgg_b200 import ops
    feats = _features()
    layers = ops.vgg_layers(feats)
    x = torch.rand(1, 3, 32, 32, device='cuda') * 3e6          # relu(conv1_1) overflows fp16
    ops._lib.load().sgg_conv_overflow(1)
    ops.vgg_features(x, layers[:2])
    assert ops._lib.load().sgg_conv_overflow(1) != 0


def test_model_backbone_uses_the_tensor_core_stack_and_matches_cudnn():
    from sgg_b200.trainstep import FakeData
    from sgg_b200.model import RelModelStanford
    import os
    torch.manual_seed(0)
    with torch.device('cuda'):
        m = RelModelStanford(train_data=FakeData(), mode='predcls').eval()
    x = torch.rand(2, 3, 64, 96, device='cuda')
    with torch.no_grad():
        a = m._backbone(x)
        os.environ['SGG_BACKBONE'] = 'cudnn'
        try:
            b = m._backbone(x)
        finally:
            del os.environ['SGG_BACKBONE']
    assert m._vgg_layers and a.shape == b.shape == (2, 512, 4, 6)
    assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max()))


_VARIANT_CHECK = r'''
import sys, torch, torch.nn as nn
sys.path.insert(0, %r)
from sgg_b200 import ops
from sgg_b200.model import _vgg16_parts
torch.manual_seed(0)
feats, _ = _vgg16_parts()
feats = feats.cuda().eval()
layers = ops.vgg_layers(feats)
x = torch.rand(3, 3, 48, 80, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))
with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
    for n_conv, upto in ((4, 10), (13, len(list(feats.children())))):     # a pooled prefix (odd tile count) and the whole stack
        r = nn.Sequential(*list(feats.children())[:upto])(x)
        y = ops.vgg_features(x, layers[:n_conv])
        err = float((y - r).abs().max()) / float(r.abs().max())
        assert y.shape == r.shape and err <= 1e-4, (n_conv, err)
assert ops._lib.load().sgg_conv_overflow(1) == 0
print('ok')
'''


@pytest.mark.parametrize('env', [{'SGG_CONV_V': '1', 'SGG_CONV_CG': '1'}, {'SGG_CONV_V': '1', 'SGG_CONV_CG': '2'},
                                 {'SGG_CONV_V': '2', 'SGG_CONV_CG': '1'}, {'SGG_CONV_V': '2', 'SGG_CONV_CG': '2'}])
def test_conv_kernel_variants_vs_cudnn_fp32(env):
    """Every kernel stays exact on every layer (the default picks v1 or v2 per layer): v1 (per-tap boxes, one tile per CTA),
    v2 (slabs, persistent CTAs) and the cta_group::2 pairs of both.
    The variant is chosen once per process (environment), hence the subprocess."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, '-c', _VARIANT_CHECK % root], env=e, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), (out.stdout[-400:], out.stderr[-800:])
