"""Names imported by the reference's utilities module for the channel solvers (out of scope)."""
import numpy as np


def aligned(shape, n=32, dtype=float, fill=None):
    a = np.empty(shape, dtype=dtype)
    if fill is not None:
        a.fill(fill)
    return a


def aligned_like(z, fill=None):
    return aligned(z.shape, dtype=z.dtype, fill=fill)


def dctn(*args, **kwargs):
    raise NotImplementedError('channel solvers are outside the B200 hot path')
