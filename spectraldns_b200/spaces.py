"""Function spaces and arrays with the surface the reference's solver modules and demos use from
shenfun, backed by the CUDA plan instead of FFTW/MPI.

Host arrays stay numpy (user callbacks read and write them, demo/Isotropic.py:161-187); every
transform copies its input to the GPU, runs the sm_100a passes through the C ABI and copies the
result back.  The time-stepping hot path does NOT go through here: integrate() keeps the state
resident on the device (see compat/spectralDNS/solvers/_device.py).

Call sites this serves: solvers/NS.py:17-64,86-110; MHD.py:18-66,79-87; tests/TG.py:24-36,100-102;
demo/Isotropic.py:33-76,159-254.
"""
import numpy as np

from .plan import Plan
from . import _lib


def world():
    """(rank, nranks, cuda device) of this process: torch.distributed when a process group exists
    (one process per GPU, launched by torchrun), else a single rank on the current device."""
    import torch
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), torch.cuda.current_device()
    except Exception:
        pass
    return 0, 1, torch.cuda.current_device() if torch.cuda.is_available() else 0


class FunctionSpace(object):
    """FunctionSpace(N, 'F', domain=(0, L), dtype=...) (solvers/NS.py:17-19)."""
    def __init__(self, N, family='F', domain=(0, 2*np.pi), dtype=float, **kw):
        if str(family).upper() not in ('F', 'FOURIER'):
            raise NotImplementedError('only Fourier bases are on the B200 path')
        self.N = int(N)
        self.domain = (float(domain[0]), float(domain[1]))
        self.dtype = np.dtype(dtype)

    def family(self):
        return 'fourier'


class Engine(object):
    """One C plan (spaces T and Tp of one grid) plus reusable device staging buffers."""
    _cache = {}

    def __init__(self, N, L, precision, dealias, solver='NS', mask_nyquist=True, decomposition='slab',
                 convection=None):
        self.key = (tuple(int(n) for n in N), tuple(float(l) for l in L), precision, dealias, solver,
                    bool(mask_nyquist), decomposition, convection)
        rank, nranks, device = world()
        self.rank, self.nranks = rank, nranks
        self.plan = Plan(self.key[0], self.key[1], precision, dealias, solver,
                         convection=convection, mask_nyquist=mask_nyquist,
                         decomposition='slab' if nranks > 1 else decomposition,
                         device=device, rank=rank, nranks=nranks)
        self._stage = {}

    @classmethod
    def get(cls, N, L, precision, dealias, solver='NS', mask_nyquist=True, decomposition='slab', convection=None):
        key = (tuple(int(n) for n in N), tuple(float(l) for l in L), precision, dealias, solver,
               bool(mask_nyquist), decomposition, convection)
        e = cls._cache.get(key)
        if e is None:
            e = cls(N, L, precision, dealias, solver, mask_nyquist, decomposition, convection)
            cls._cache[key] = e
        return e

    def stage(self, tag, shape, dtype):
        import torch
        k = (tag, tuple(shape), dtype)
        t = self._stage.get(k)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.plan.device)
            self._stage[k] = t
        return t

    def upload(self, tag, a, dtype_np, tdtype):
        import torch
        a = np.ascontiguousarray(a, dtype=dtype_np)
        t = self.stage(tag, a.shape, tdtype)
        t.copy_(torch.from_numpy(a))
        return t


class TensorProductSpace(object):
    """3-D r2c Fourier space (solvers/NS.py:21-25).  `which` selects the plan's plain space T or
    its dealiased companion Tp (NS.py:29-31)."""

    def __init__(self, comm, bases, dtype=None, slab=True, collapse_fourier=True,
                 padding_factor=1, dealias_direct=False, engine=None, which=None, **kw):
        self.comm = comm
        self.bases = list(bases)
        self.N = tuple(b.N for b in self.bases)
        self.L = tuple(b.domain[1]-b.domain[0] for b in self.bases)
        self.float = np.dtype(dtype if dtype is not None else float)
        self.complex = np.dtype(np.complex64 if self.float == np.float32 else np.complex128)
        self.slab = slab
        self.padding_factor = padding_factor
        self.dealias_direct = dealias_direct
        padded = padding_factor > 1.0 + 1e-8
        self.M = tuple(int(np.floor(n*padding_factor)) for n in self.N) if padded else self.N
        self._engine = engine
        self._which = which if which is not None else (
            _lib.SPACE_TP if (padded or dealias_direct) else _lib.SPACE_T)
        self._solver = kw.pop('solver', 'NS')
        self._mask_nyquist = kw.pop('mask_nyquist', True)

    # -- engine -----------------------------------------------------------
    def _dealias_name(self):
        if self.padding_factor > 1.0 + 1e-8:
            return '3/2-rule'
        return '2/3-rule' if self.dealias_direct else 'None'

    @property
    def engine(self):
        if self._engine is None:
            prec = 'single' if self.float == np.float32 else 'double'
            self._engine = Engine.get(self.N, self.L, prec, self._dealias_name(), self._solver,
                                      self._mask_nyquist, 'slab' if self.slab else 'pencil')
        return self._engine

    # -- shapes / meshes (NS.py:37-48) ---------------------------------------
    def _ranks(self):
        if self._engine is not None:
            return self._engine.rank, self._engine.nranks
        r, n, _ = world()
        return r, n

    def global_shape(self, forward_output=False):
        if forward_output:
            return (self.N[0], self.N[1], self.N[2]//2+1)
        return tuple(self.M)

    def shape(self, forward_output=False):
        """Local shape: slab decomposition, spectral space split on axis 1, physical on axis 0
        (spectralDNS3D_short.py:28-29)."""
        r, n = self._ranks()
        g = self.global_shape(forward_output)
        if n == 1:
            return g
        return (g[0], g[1]//n, g[2]) if forward_output else (g[0]//n, g[1], g[2])

    def local_slice(self, forward_output=False):
        r, n = self._ranks()
        g = self.global_shape(forward_output)
        ax = 1 if forward_output else 0
        out = [slice(0, m) for m in g]
        c = g[ax]//n
        out[ax] = slice(r*c, (r+1)*c)
        return tuple(out)

    def dims(self):
        return 3

    def __len__(self):
        return 3

    def local_mesh(self, broadcast=False):
        X = []
        for i in range(3):
            s = [1, 1, 1]
            s[i] = self.M[i]
            x = (np.arange(self.M[i], dtype=float)*self.L[i]/self.M[i])
            x = x[self.local_slice(False)[i]]
            s[i] = len(x)
            x = x.reshape(s)
            X.append(np.broadcast_to(x, self.shape(False)) if broadcast else x)
        return X

    def local_wavenumbers(self, broadcast=False, scaled=False, eliminate_highest_freq=False):
        K = []
        for i in range(3):
            n = self.N[i]
            k = np.fft.fftfreq(n, 1./n) if i < 2 else np.fft.rfftfreq(n, 1./n)
            if scaled:
                k = k*2*np.pi/self.L[i]
            k = k[self.local_slice(True)[i]]
            s = [1, 1, 1]
            s[i] = len(k)
            k = k.reshape(s)
            K.append(np.broadcast_to(k, self.shape(True)) if broadcast else k)
        return K

    def get_mask_nyquist(self):
        mask = np.ones(self.global_shape(True), dtype=int)
        for i, n in enumerate(self.N):
            if n % 2 == 0:
                s = [slice(None)]*3
                s[i] = n//2
                mask[tuple(s)] = 0
        return np.ascontiguousarray(mask[self.local_slice(True)])

    def mask_nyquist(self, u_hat, mask=None):
        u_hat *= (self.get_mask_nyquist() if mask is None else mask)
        return u_hat

    def get_dealiased(self, padding_factor=1.5, dealias_direct=False):
        padded = padding_factor > 1.0 + 1e-8
        name = '3/2-rule' if padded else ('2/3-rule' if dealias_direct else 'None')
        eng = self._engine
        if eng is not None and eng.key[3] != name:
            eng = None
        return TensorProductSpace(self.comm, self.bases, dtype=self.float, slab=self.slab,
                                  padding_factor=padding_factor, dealias_direct=dealias_direct,
                                  engine=eng, which=_lib.SPACE_TP if (padded or dealias_direct) else _lib.SPACE_T,
                                  solver=self._solver, mask_nyquist=self._mask_nyquist)

    # -- transforms ---------------------------------------------------------
    def _run(self, forward, src, dst, ncomp):
        eng = self.engine
        p = eng.plan
        tp = self._which == _lib.SPACE_TP
        if forward:
            d_in = eng.upload('fi', np.asarray(src).reshape((ncomp,)+self.shape(False)), p.float, p.tfloat)
            d_out = eng.stage('fo', (ncomp,)+self.shape(True), p.tcomplex)
            p.use_current_stream()
            p.forward(d_in, out=d_out, padded=tp)
        else:
            d_in = eng.upload('bi', np.asarray(src).reshape((ncomp,)+self.shape(True)), p.complex, p.tcomplex)
            d_out = eng.stage('bo', (ncomp,)+self.shape(False), p.tfloat)
            p.use_current_stream()
            p.backward(d_in, out=d_out, padded=tp)
        res = d_out.cpu().numpy().reshape(np.shape(dst))
        np.copyto(dst, res, casting='same_kind')
        return dst

    def forward(self, input_array, output_array=None):
        if output_array is None:
            output_array = Function(self)
        return self._run(True, input_array, output_array, 1)

    def backward(self, input_array, output_array=None):
        if output_array is None:
            output_array = Array(self)
        return self._run(False, input_array, output_array, 1)


class CompositeSpace(object):
    """CompositeSpace([T]*n) / VectorSpace(T): leading component axis (NS.py:26,32; MHD.py:27,36)."""
    def __init__(self, spaces):
        self.spaces = list(spaces)
        self.T = self.spaces[0]
        self.ncomp = len(self.spaces)

    def shape(self, forward_output=False):
        return (self.ncomp,) + self.T.shape(forward_output)

    global_shape = shape

    def local_slice(self, forward_output=False):
        return (slice(0, self.ncomp),) + self.T.local_slice(forward_output)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return self.ncomp

    def forward(self, input_array, output_array=None):
        if output_array is None:
            output_array = Function(self)
        n = int(np.shape(input_array)[0])
        return self.T._run(True, input_array, output_array, n)

    def backward(self, input_array, output_array=None):
        if output_array is None:
            output_array = Array(self)
        n = int(np.shape(input_array)[0])
        return self.T._run(False, input_array, output_array, n)

    def __getattr__(self, name):
        if name in ('spaces', 'T', 'ncomp'):
            raise AttributeError(name)
        return getattr(self.T, name)


class VectorSpace(CompositeSpace):
    def __init__(self, T):
        CompositeSpace.__init__(self, [T]*3)


def _scalar_space(space):
    return space.T if isinstance(space, CompositeSpace) else space


class _SpaceArray(np.ndarray):
    """ndarray that remembers its function space; slices and views keep it (UB[:3], U_hat[0])."""
    _spectral = False

    def __new__(cls, space, val=0, buffer=None, **kw):
        shape = space.shape(cls._spectral)
        T = _scalar_space(space)
        dtype = T.complex if cls._spectral else T.float
        if buffer is not None:
            base = np.asarray(buffer)
            if base.dtype != dtype:
                base = base.view(dtype)
            obj = base.reshape(shape).view(cls)
        else:
            obj = np.zeros(shape, dtype=dtype).view(cls)
            if val != 0:
                obj.fill(val)
        obj._space = space
        return obj

    def __array_finalize__(self, obj):
        self._space = getattr(obj, '_space', None)

    def function_space(self):
        return self._space

    def _transform_space(self):
        """Space whose component count matches this (possibly sliced) array."""
        T = _scalar_space(self._space)
        return T if self.ndim == 3 else CompositeSpace([T]*self.shape[0])


class Array(_SpaceArray):
    """Physical-space array (NS.py:51,55)."""
    _spectral = False

    def forward(self, output_array=None):
        return self._transform_space().forward(self, output_array)


class Function(_SpaceArray):
    """Spectral-space array (NS.py:52,54,61-63)."""
    _spectral = True

    def backward(self, output_array=None):
        return self._transform_space().backward(self, output_array)

    def mask_nyquist(self, mask=None):
        _scalar_space(self._space).mask_nyquist(self, mask)
        return self


class CachedArrayDict(dict):
    """work[(like, idx, zero)]: cached scratch array shaped/typed like `like`, one per idx,
    zero-filled on fetch when `zero` (solvers/NS.py:126,133,194; VV.py:64,93-94)."""
    def __getitem__(self, key):
        like, idx, zero = key
        if isinstance(like, tuple):
            shape, dtype = tuple(like[0]), np.dtype(like[1])
        else:
            shape, dtype = like.shape, like.dtype
        k = (shape, dtype.str, idx)
        if not dict.__contains__(self, k):
            dict.__setitem__(self, k, np.zeros(shape, dtype=dtype))
        a = dict.__getitem__(self, k)
        if zero:
            a.fill(0)
        return a


def energy_fourier(u_hat, T):
    """Hermitian-weighted sum |u_hat|^2 (tests/TG.py:101; demo/Isotropic.py:67,167-182).
    Host array in, float out; the reduction runs on the GPU (sdns_energy)."""
    S = _scalar_space(T)
    a = np.asarray(u_hat)
    eng = S.engine
    p = eng.plan
    nc = int(a.size // int(np.prod(S.shape(True))))
    d = eng.upload('en', a.reshape((nc,)+S.shape(True)), p.complex, p.tcomplex)
    p.use_current_stream()
    res = p.energy(d)
    comm = getattr(S, 'comm', None)
    if comm is not None and hasattr(comm, 'allreduce'):
        res = comm.allreduce(res)
    return float(res)
