"""One VGG conv layer in a loop (ncu target / clock sampling): CL = layer index (1..12), CB = batch, CREPS = launches, CSECS = min seconds."""
import os, sys, time, subprocess, threading
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgg_b200 import ops, _lib
from sgg_b200.ops import _ptr, _stream, check
from sgg_b200.model import _vgg16_parts
torch.cuda.set_device(0)
torch.manual_seed(0)
B, L = int(os.environ.get('CB', 32)), int(os.environ.get('CL', 8))
feats, _ = _vgg16_parts()
layers = ops.vgg_layers(feats.cuda().eval())
h = w = 608; cin = 64
for li, (wt, bs, pool) in enumerate(layers[1:], start=1):
    cout = wt.shape[0]
    if li == L:
        break
    if pool: h //= 2; w //= 2
    cin = cout
x = (torch.rand((2, B, h, w, cin), device='cuda') * 0.1).half()
wp = ops.conv_weight_planes(wt)
ho, wo = (h // 2, w // 2) if pool else (h, w)
out = torch.empty((2, B, ho, wo, cout), dtype=torch.float16, device='cuda')
lib = _lib.load()
def run():
    check(lib.sgg_conv3x3_tc(_ptr(x), _ptr(wp), _ptr(bs), B, h, w, cin, cout, 1, 1 if pool else 0, _ptr(out), None, _stream()), 'conv')
reps = int(os.environ.get('CREPS', 3))
secs = float(os.environ.get('CSECS', 0))
samples = []
stop = False
def sampler():
    while not stop:
        try:
            o = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active', '--format=csv,noheader,nounits', '-i', '0'],
                               capture_output=True, text=True, timeout=5).stdout.strip()
            samples.append(o)
        except Exception as e:
            samples.append(str(e))
        time.sleep(0.2)
if secs > 0:
    th = threading.Thread(target=sampler); th.start()
    t0 = time.time(); n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.time() - t0 < secs:
        for _ in range(20): run()
        n += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    stop = True; th.join()
    print('layer %d: %d launches, %.3f ms each (incl. host sync gaps)' % (L, n, a.elapsed_time(b) / n))
    print('clock samples (sm MHz, mem MHz, W, reasons):', samples[2:-1][:12])
else:
    for _ in range(reps): run()
    torch.cuda.synchronize()
    print('done layer', L, (cin, cout, h, w))
if os.environ.get('CEACH'):
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.time()
        a.record(); run(); b.record(); torch.cuda.synchronize(); t1 = time.time()
        print('launch %d: events %.3f ms, wall %.3f ms' % (i, a.elapsed_time(b), (t1 - t0) * 1e3))
