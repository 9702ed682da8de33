"""CPU: `bench.py --impl reference` (the reference's CPU path through the oracle port) prints ONE JSON line carrying the
contract keys, with the same `metric` / `unit` / `config.workload` the native arm reports, and needs no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True
    assert d['metric'] == 'images/sec (PredCls, 3 MP iters)' and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['value'] > 0 and abs(d['value'] - 8.0 / (d['ms_per_step'] * 1e-3)) <= 1e-6 * d['value']
    assert d['config']['workload'].startswith('PredCls L1 batch=8, 30 boxes/img, 300 edges/img')
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == d['value']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_reference_arm_under_torchrun_prints_once():
    """Launched like the driver launches N > 1 (torchrun, one process per rank): rank 0 alone runs and prints the line,
    the other rank exits 0 without work."""
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29519', os.path.join(ROOT, 'bench.py'),
                          '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0


def test_native_arm_fails_loudly_without_a_gpu():
    """No CPU fallback on the product path: without a CUDA device the native arm exits non-zero and prints no JSON line."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('a CUDA device is present')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert not [ln for ln in out.stdout.splitlines() if ln.strip().startswith('{')]
    assert 'CUDA' in out.stderr
