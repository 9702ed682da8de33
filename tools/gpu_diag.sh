#!/bin/bash
# Environment + timing diagnostics of the GPU box (why CPU-side stages are slow there).
mkdir -p gpurun_out
{
echo "nproc=$(nproc) getconf=$(getconf _NPROCESSORS_ONLN)"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /proc/loadavg
echo "OMP_NUM_THREADS=$OMP_NUM_THREADS MKL_NUM_THREADS=$MKL_NUM_THREADS"; free -g | head -2
grep -m1 "model name" /proc/cpuinfo
} > gpurun_out/diag_env.txt 2>&1
cat gpurun_out/diag_env.txt
timeout 200 python - > gpurun_out/diag_py.txt 2>&1 <<'PY'
import time, os
t0 = time.time()
def lap(msg):
    print('%7.2fs %s' % (time.time() - t0, msg), flush=True)
lap('start; cpu_count=%s affinity=%d' % (os.cpu_count(), len(os.sched_getaffinity(0))))
import numpy as np
lap('numpy')
import torch
lap('torch imported; threads=%d interop=%d' % (torch.get_num_threads(), torch.get_num_interop_threads()))
a = torch.randn(4096, 4096)
lap('cpu randn 16M')
b = a @ a
lap('cpu matmul 4096^3 default threads')
torch.set_num_threads(8)
b = a @ a
lap('cpu matmul 4096^3 8 threads')
b = a @ a
lap('cpu matmul 4096^3 8 threads (2nd)')
x = np.random.default_rng(0).standard_normal((4096, 25088), dtype=np.float32)
lap('numpy 100M normals')
torch.cuda.init(); torch.zeros(1, device='cuda'); torch.cuda.synchronize()
lap('cuda init')
import sys; sys.path.insert(0, '.')
from sgg_b200 import ops, _lib
lib = _lib.load()
lap('lib loaded, engine %s' % ops.tc_engine())
xg = torch.randn(1000, 4096, device='cuda'); wg = torch.randn(4096, 4096, device='cuda') / 64
torch.cuda.synchronize(); lap('gpu randn')
for mode in ('simt', 'tc32', 'tc16'):
    ops.set_gemm_mode(mode)
    y = ops.linear(xg, wg); torch.cuda.synchronize()
    lap('linear 1000x4096x4096 %s first' % mode)
    for _ in range(5): y = ops.linear(xg, wg)
    torch.cuda.synchronize()
    lap('linear x5 %s' % mode)
ref = xg.double() @ wg.double().t(); torch.cuda.synchronize()
lap('fp64 ref; err %.2e' % float((y.double() - ref).abs().max()))
with torch.device('cuda'):
    from sgg_b200.model import RelModelStanford
    class FakeData:
        ind_to_classes = ['__background__'] + ['c%d' % i for i in range(150)]
        ind_to_predicates = ['__background__'] + ['p%d' % i for i in range(50)]
    m = RelModelStanford(train_data=FakeData(), mode='predcls')
torch.cuda.synchronize()
lap('model constructed on cuda')
PY
cat gpurun_out/diag_py.txt
echo "=== tc16 probe (mp only)"; SGG_TC_MODE=1 SGG_CHECK_SKIP_LINEAR=1 timeout 150 python tools/tc16_check.py > gpurun_out/c2_tc16_check.log 2>&1; echo "rc=$?"; tail -20 gpurun_out/c2_tc16_check.log
echo "=== pytest parity (durations)"; OMP_NUM_THREADS=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=15 > gpurun_out/diag_pytest.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/diag_pytest.log
echo "=== pytest model (durations)"; timeout 400 python -m pytest tests/test_gpu_model.py -m gpu -x -q --durations=15 > gpurun_out/diag_pytest_model.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/diag_pytest_model.log
