#!/usr/bin/env python
"""Build the reference's ONLY native component from its own source, where it lies.

    /root/reference/lib/draw_rectangles/draw_rectangles.pyx  (Cython, 67 lines)
        -> oracle/_ref/draw_rectangles*.so

Recipe: cythonize the .pyx straight from the read-only reference tree into
oracle/_ref/ (C file and .so both land there; nothing is copied into the repo's
history — oracle/_ref/ is git-ignored but travels to the GPU box).  The checked-in
draw_rectangles.c of the reference is Cython-0.29 output that does not build on
Python 3.12, so it is regenerated, not used.  Everything else on the hot path is
Python/PyTorch and cannot be "compiled"; it is run in the build container by
tests/golden/make_golden.py instead.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
SRC = '/root/reference/lib/draw_rectangles/draw_rectangles.pyx'


def build(verbose=False):
    if not os.path.exists(SRC):
        return None                      # GPU box: only the prebuilt files (if any) are used
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, 'draw_rectangles' + sysconfig.get_config_var('EXT_SUFFIX'))
    if os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(SRC):
        return so
    import numpy as np
    c_file = os.path.join(OUT, 'draw_rectangles.c')
    subprocess.check_call([sys.executable, '-m', 'cython', '-3', SRC, '-o', c_file])
    inc = sysconfig.get_paths()['include']
    cmd = ['gcc', '-O2', '-shared', '-fPIC', '-w', '-I', inc, '-I', np.get_include(), c_file, '-o', so]
    subprocess.check_call(cmd)
    if verbose:
        print('built', so)
    return so


def load():
    """Import the built module (None if it is not available)."""
    import importlib.util
    if not os.path.isdir(OUT):
        return None
    for f in os.listdir(OUT):
        if f.startswith('draw_rectangles') and f.endswith('.so'):
            spec = importlib.util.spec_from_file_location('draw_rectangles', os.path.join(OUT, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


if __name__ == '__main__':
    print(build(verbose=True))
