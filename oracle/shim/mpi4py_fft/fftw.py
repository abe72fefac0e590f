"""TEST INFRASTRUCTURE ONLY -- names imported by spectralDNS/utilities/__init__.py:13 (channel only)."""
import numpy as np


def aligned(shape, n=32, dtype=float, fill=None):
    a = np.empty(shape, dtype=dtype)
    if fill is not None:
        a.fill(fill)
    return a


def aligned_like(z, fill=None):
    return aligned(z.shape, dtype=z.dtype, fill=fill)


def dctn(*a, **k):
    raise NotImplementedError
