"""TEST INFRASTRUCTURE ONLY -- numpy/scipy.fft stand-in for the `shenfun` names that
the reference's triply periodic solvers touch.

The reference (spectralDNS v1.4.0) imports shenfun, which is absent from this image and
whose source is not under /root/reference.  This module restates, in plain numpy, the
minimal behavioural contract of shenfun *as the reference uses it* (SURVEY.md Appendix A),
so that the reference's UNMODIFIED solvers/NS.py, VV.py, MHD.py, maths/integrators.py and
tests/TG.py, tests/TGMHD.py can be imported from /root/reference in the build container to
produce golden vectors (oracle/make_golden.py).  Single rank only.

Conventions (each pinned by a reference call site):
  * forward  = rfftn(u) / prod(M)  (tests/TG.py:98-109: energy_fourier(U_hat)/2 == sum(U*U)/prod(N)/2)
  * backward = irfftn(u_hat) * prod(M)
  * wavenumbers: fftfreq order on axes 0,1; rfftfreq on axis 2 with Nyquist +N/2, scaled by
    2*pi/L, returned as broadcastable (N0,1,1),(1,N1,1),(1,1,Nh) arrays
    (optimization/cython_solvers.in:41, cython_maths.in:68-76)
  * mask_nyquist zeroes every mode with an index N_i/2 (solvers/NS.py:34,253-254)
  * get_dealiased(padding_factor=1, dealias_direct=True): truncation applied to the INPUT of
    backward (solvers/NS.py:29-31) with the in-tree cutoff |k_i| < 2/3*(N_i/2+1)
    (spectralDNS3D_short.py:44-46)
  * get_dealiased(padding_factor=1.5): per-axis zero padding N -> floor(1.5 N) on backward and
    corner truncation on forward (solvers/NS.py:15,29-31)

Nothing in the product package may import this module.
"""
import numpy as np
import scipy.fft as sfft

__all__ = ['FunctionSpace', 'TensorProductSpace', 'VectorSpace', 'CompositeSpace',
           'Array', 'Function', 'CachedArrayDict', 'ShenfunFile']

WORKERS = -1


class FunctionSpace(object):
    def __init__(self, N, family='F', domain=(0, 2*np.pi), dtype=float, **kw):
        assert family.upper() in ('F', 'FOURIER')
        self.N = int(N)
        self.domain = (float(domain[0]), float(domain[1]))
        self.dtype = np.dtype(dtype)


def dealias_cutoff(N):
    """Largest kept |k| (integer wavenumber) for the 2/3-rule: |k| < 2/3*(N/2+1)
    (spectralDNS3D_short.py:44-46)."""
    kmax = 2./3.*(N//2+1)
    kc = int(np.ceil(kmax)) - 1
    return kc


class TensorProductSpace(object):
    def __init__(self, comm, bases, dtype=None, slab=True, collapse_fourier=True,
                 padding_factor=1, dealias_direct=False, **kw):
        self.comm = comm
        self.bases = list(bases)
        self.N = tuple(b.N for b in bases)
        self.L = tuple(b.domain[1]-b.domain[0] for b in bases)
        self.float = np.dtype(dtype if dtype is not None else float)
        self.complex = np.dtype(np.complex64 if self.float == np.float32 else np.complex128)
        self.padding_factor = padding_factor
        self.dealias_direct = dealias_direct
        self.M = tuple(int(np.floor(n*padding_factor)) for n in self.N)
        self.slab = slab
        self.rank_ = 0  # tensor rank (scalar space)

    # -- shapes -----------------------------------------------------------
    def shape(self, forward_output=False):
        if forward_output:
            return (self.N[0], self.N[1], self.N[2]//2+1)
        return tuple(self.M)

    def global_shape(self, forward_output=False):
        return self.shape(forward_output)

    def local_slice(self, forward_output=False):
        return tuple(slice(0, n) for n in self.shape(forward_output))

    def dims(self):
        return 3

    def local_mesh(self, broadcast=False):
        X = []
        for i in range(3):
            x = np.arange(self.M[i], dtype=float)*self.L[i]/self.M[i]
            s = [1, 1, 1]
            s[i] = self.M[i]
            x = x.reshape(s)
            if broadcast:
                x = np.broadcast_to(x, self.M)
            X.append(x)
        return X

    def local_wavenumbers(self, broadcast=False, scaled=False, eliminate_highest_freq=False):
        K = []
        for i in range(3):
            n = self.N[i]
            if i < 2:
                k = np.fft.fftfreq(n, 1./n)
            else:
                k = np.fft.rfftfreq(n, 1./n)
            if scaled:
                k = k*2*np.pi/self.L[i]
            s = [1, 1, 1]
            s[i] = len(k)
            k = k.reshape(s)
            if broadcast:
                k = np.broadcast_to(k, self.shape(True))
            K.append(k)
        return K

    # -- masks ------------------------------------------------------------
    def get_mask_nyquist(self):
        mask = np.ones(self.shape(True), dtype=int)
        for i in range(3):
            n = self.N[i]
            if n % 2 == 0:
                s = [slice(None)]*3
                s[i] = n//2
                mask[tuple(s)] = 0
        return mask

    def mask_nyquist(self, u_hat, mask=None):
        if mask is None:
            mask = self.get_mask_nyquist()
        u_hat *= mask
        return u_hat

    def _dealias_mask(self):
        if not hasattr(self, '_dmask'):
            m = np.ones(self.shape(True), dtype=bool)
            K = self.local_wavenumbers(scaled=False)
            for i in range(3):
                kc = dealias_cutoff(self.N[i])
                m = m & (np.abs(K[i]) <= kc)
            self._dmask = m
        return self._dmask

    def get_dealiased(self, padding_factor=1.5, dealias_direct=False):
        return TensorProductSpace(self.comm, self.bases, dtype=self.float, slab=self.slab,
                                  padding_factor=padding_factor, dealias_direct=dealias_direct)

    # -- transforms -------------------------------------------------------
    def forward(self, u, u_hat=None):
        M, N = self.M, self.N
        full = sfft.rfftn(np.asarray(u), axes=(0, 1, 2), workers=WORKERS)
        full = full/np.prod(M)
        if M != N:
            out = np.zeros(self.shape(True), dtype=full.dtype)
            n0, n1, nh = N[0], N[1], N[2]//2+1
            h0, h1 = n0//2, n1//2
            # low/high corner blocks of axes 0,1; first Nh of axis 2
            out[:h0, :h1] = full[:h0, :h1, :nh]
            out[:h0, h1:] = full[:h0, M[1]-(n1-h1):, :nh]
            out[h0:, :h1] = full[M[0]-(n0-h0):, :h1, :nh]
            out[h0:, h1:] = full[M[0]-(n0-h0):, M[1]-(n1-h1):, :nh]
            full = out
        if u_hat is None:
            u_hat = Function(self)
        u_hat[...] = full
        return u_hat

    def backward(self, u_hat, u=None):
        M, N = self.M, self.N
        a = np.asarray(u_hat)
        if M != N:
            n0, n1, nh = N[0], N[1], N[2]//2+1
            h0, h1 = n0//2, n1//2
            full = np.zeros((M[0], M[1], M[2]//2+1), dtype=a.dtype)
            full[:h0, :h1, :nh] = a[:h0, :h1]
            full[:h0, M[1]-(n1-h1):, :nh] = a[:h0, h1:]
            full[M[0]-(n0-h0):, :h1, :nh] = a[h0:, :h1]
            full[M[0]-(n0-h0):, M[1]-(n1-h1):, :nh] = a[h0:, h1:]
            a = full
        elif self.dealias_direct:
            a = a*self._dealias_mask()
        r = sfft.irfftn(a, s=M, axes=(0, 1, 2), workers=WORKERS)*np.prod(M)
        if u is None:
            u = Array(self)
        u[...] = r
        return u


class CompositeSpace(object):
    def __init__(self, spaces):
        self.spaces = list(spaces)
        self.T = spaces[0]
        self.ncomp = len(spaces)

    def shape(self, forward_output=False):
        return (self.ncomp,) + self.T.shape(forward_output)

    def local_slice(self, forward_output=False):
        return (slice(0, self.ncomp),) + self.T.local_slice(forward_output)

    def __getitem__(self, i):
        return self.spaces[i]

    def forward(self, u, u_hat=None):
        if u_hat is None:
            u_hat = Function(self)
        for i in range(self.ncomp):
            self.T.forward(u[i], u_hat[i])
        return u_hat

    def backward(self, u_hat, u=None):
        if u is None:
            u = Array(self)
        for i in range(self.ncomp):
            self.T.backward(u_hat[i], u[i])
        return u

    def __getattr__(self, name):
        if name in ('spaces', 'T', 'ncomp'):
            raise AttributeError(name)
        return getattr(self.T, name)


class VectorSpace(CompositeSpace):
    def __init__(self, T):
        CompositeSpace.__init__(self, [T]*3)


class _SpaceArray(np.ndarray):
    _forward_output = False

    def __new__(cls, space, val=0, buffer=None, **kw):
        shape = space.shape(cls._forward_output)
        T = space.T if isinstance(space, CompositeSpace) else space
        dtype = T.complex if cls._forward_output else T.float
        if buffer is not None:
            obj = np.frombuffer(buffer, dtype=dtype, count=int(np.prod(shape))).reshape(shape) \
                if not isinstance(buffer, np.ndarray) else buffer.reshape(shape)
            obj = obj.view(cls)
        else:
            obj = np.ndarray.__new__(cls, shape, dtype=dtype)
            obj.fill(val)
        obj._space = space
        return obj

    def __array_finalize__(self, obj):
        self._space = getattr(obj, '_space', None)

    def function_space(self):
        return self._space


class Array(_SpaceArray):
    _forward_output = False

    def forward(self, output_array=None):
        return self._space.forward(self, output_array)


class Function(_SpaceArray):
    _forward_output = True

    def backward(self, output_array=None):
        return self._space.backward(self, output_array)

    def mask_nyquist(self, mask=None):
        T = self._space.T if isinstance(self._space, CompositeSpace) else self._space
        T.mask_nyquist(self, mask)
        return self


class CachedArrayDict(dict):
    """work[(like, idx, zero)] -> cached scratch array with like's shape/dtype, distinct per
    idx, zero-filled on fetch when zero is true (solvers/NS.py:126,133,140,151,194)."""
    def __getitem__(self, key):
        like, idx, zero = key
        if isinstance(like, tuple):
            shape, dtype = like[0], np.dtype(like[1])
        else:
            shape, dtype = like.shape, like.dtype
        k = (tuple(shape), dtype.str, idx)
        if not dict.__contains__(self, k):
            dict.__setitem__(self, k, np.zeros(shape, dtype=dtype))
        a = dict.__getitem__(self, k)
        if zero:
            a.fill(0)
        return a


class _Attrs(dict):
    def create(self, k, v):
        self[k] = v


class _F(object):
    def __init__(self):
        self.attrs = _Attrs()


class ShenfunFile(object):
    """No-op stand-in (h5py is absent): keeps tstep/t attrs in memory only
    (h5io/HDF5File.py:66-89)."""
    def __init__(self, name, space, mode='w', **kw):
        self.filename = name
        self._f = _F()
        self.f = None
        self.writes = 0

    def open(self):
        self.f = self._f

    def close(self):
        self.f = None

    def write(self, tstep, data, as_scalar=False):
        self.writes += 1
