"""TEST INFRASTRUCTURE ONLY -- stand-in for shenfun.fourier.energy_fourier (tests/TG.py:5,101)."""
import numpy as np


def energy_fourier(u_hat, T):
    """sum |u_hat|^2 with weight 2 on 0<k2<N2/2 and 1 on k2=0 and the Nyquist plane, so that it
    equals sum(u*u)/prod(N) of the physical field (tests/TG.py:98-109)."""
    a = np.asarray(u_hat)
    N2 = T.N[2] if hasattr(T, 'N') else T.T.N[2]
    w = (a.real.astype(np.float64)**2 + a.imag.astype(np.float64)**2)
    if N2 % 2 == 0:
        res = 2*np.sum(w[..., 1:-1]) + np.sum(w[..., 0]) + np.sum(w[..., -1])
    else:
        res = 2*np.sum(w[..., 1:]) + np.sum(w[..., 0])
    return res
