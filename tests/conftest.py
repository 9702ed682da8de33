import os, sys

# The GPU boxes expose far more logical CPUs than the container may use: an unbounded OpenMP / MKL pool makes every
# CPU-side op (parameter init, the numpy oracle) crawl.  Bound the pools before numpy / torch are imported.
for _v in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
    os.environ.setdefault(_v, '8')
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree C-ABI library; builds it (nvcc cross-compile) if missing."""
    from sgg_b200 import _build
    return _build.ensure_built()
