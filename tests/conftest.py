import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a.astype(np.complex128) - b).ravel()) /
                 max(np.linalg.norm(b.ravel()), 1e-300))
