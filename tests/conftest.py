import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def built_lib():
    """Build (if stale) and load liblemo_b200.so.  nvcc cross-compiles sm_100a without a GPU."""
    from lemo_b200 import _lib
    _lib.build()
    return _lib.lib()


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden.npz')))


def relerr(a, b):
    """max |a-b| / max |b| on torch tensors / numpy arrays."""
    import numpy as np
    a = a.detach().cpu().double().numpy() if hasattr(a, 'detach') else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if hasattr(b, 'detach') else np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
