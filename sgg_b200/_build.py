"""In-tree build of the C-ABI library ``sgg_b200/libsgg_b200.so`` with nvcc for sm_100a.

No torch extension machinery: the library has no torch / pybind symbols, it is
loaded with ctypes (``sgg_b200._lib``).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libsgg_b200.so')
OBJ_DIR = os.path.join(PKG, 'csrc', 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
          '--expt-relaxed-constexpr', '-Xptxas', '-v']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(os.path.dirname(PKG), 'include', 'sgg_b200.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), _deps_mtime()):
        return obj, ''
    cmd = [NVCC] + ARCH + CFLAGS + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in res]
    log = '\n'.join(l for _, l in res if l)
    if verbose and log:
        print(log)
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-Xcompiler', '-fPIC', '-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


def ensure_built():
    """Build if the library is missing or older than any source (only possible where nvcc exists)."""
    stale = not os.path.exists(LIB)
    if not stale and os.path.exists(NVCC):
        newest = max([os.path.getmtime(s) for s in sources()] + [_deps_mtime()])
        stale = newest > os.path.getmtime(LIB)
    if stale:
        if not os.path.exists(NVCC):
            raise RuntimeError('libsgg_b200.so is missing and nvcc is not available to build it')
        build()
    return LIB


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
