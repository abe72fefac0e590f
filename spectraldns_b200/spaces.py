"""Function spaces and arrays with the surface the reference's solver modules and demos use from
shenfun, backed by the CUDA plan instead of FFTW/MPI.

Host arrays stay numpy (user callbacks read and write them, demo/Isotropic.py:161-187); every
transform copies its input to the GPU, runs the sm_100a passes through the C ABI and copies the
result back.  The time-stepping hot path does NOT go through here: integrate() keeps the state
resident on the device (see compat/spectralDNS/solvers/_device.py).

Call sites this serves: solvers/NS.py:17-64,86-110; MHD.py:18-66,79-87; tests/TG.py:24-36,100-102;
demo/Isotropic.py:33-76,159-254.
"""
import os
import numpy as np

from .plan import Plan, Plan2D
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
        if len(self.key[0]) == 2:
            # doubly periodic solvers (NS2D / Bq2D): a thousandth of a 3-D problem, single GPU
            if nranks > 1:
                raise NotImplementedError('the 2-D solvers run on one GPU (a 2-D grid does not fill one)')
            self.plan = Plan2D(self.key[0], self.key[1], precision, dealias, solver if solver in ('NS2D', 'Bq2D') else 'NS2D',
                               mask_nyquist=mask_nyquist, device=device)
        else:
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
        if _LAZY and isinstance(a, np.ndarray):
            d = _lazy_dev(a)
            if d is not None:
                d.sync_to_host()            # lazily mirrored state: bring the host copy up to date before reading it
        a = np.ascontiguousarray(np.asarray(a).view(np.ndarray) if isinstance(a, np.ndarray) else a, dtype=dtype_np)
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

    @property
    def dim(self):
        return len(self.N)

    def global_shape(self, forward_output=False):
        if forward_output:
            return tuple(self.N[:-1]) + (self.N[-1]//2+1,)
        return tuple(self.M)

    def shape(self, forward_output=False):
        """Local shape: slab decomposition, spectral space split on axis 1, physical on axis 0
        (spectralDNS3D_short.py:28-29)."""
        r, n = self._ranks()
        g = self.global_shape(forward_output)
        if n == 1 or self.dim == 2:
            return g
        sl = self.local_slice(forward_output)
        return tuple(len(range(*s.indices(m))) for s, m in zip(sl, g))

    def local_slice(self, forward_output=False):
        r, n = self._ranks()
        g = self.global_shape(forward_output)
        out = [slice(0, m) for m in g]
        if self.dim == 3:
            # mpi4py-fft's slabs: m // n entries per rank, the first m % n ranks one more
            ax = 1 if forward_output else 0
            c, rem = divmod(g[ax], n)
            lo = r*c + min(r, rem)
            out[ax] = slice(lo, lo + c + (1 if r < rem else 0))
            if forward_output and n > 1 and os.environ.get('SDNS_K1_LAYOUT', 'blocks') == 'cyclic':
                out[ax] = slice(r, g[ax], n)       # Plan(k1_layout='cyclic'): rank r owns the axis-1 modes r, r + n, ...
        return tuple(out)

    def dims(self):
        return self.dim

    def __len__(self):
        return self.dim

    def local_mesh(self, broadcast=False):
        X = []
        for i in range(self.dim):
            s = [1]*self.dim
            s[i] = self.M[i]
            x = (np.arange(self.M[i], dtype=float)*self.L[i]/self.M[i])
            x = x[self.local_slice(False)[i]]
            s[i] = len(x)
            x = x.reshape(s)
            X.append(np.broadcast_to(x, self.shape(False)) if broadcast else x)
        return X

    def local_wavenumbers(self, broadcast=False, scaled=False, eliminate_highest_freq=False):
        K = []
        for i in range(self.dim):
            n = self.N[i]
            k = np.fft.fftfreq(n, 1./n) if i < self.dim-1 else np.fft.rfftfreq(n, 1./n)
            if scaled:
                k = k*2*np.pi/self.L[i]
            k = k[self.local_slice(True)[i]]
            s = [1]*self.dim
            s[i] = len(k)
            k = k.reshape(s)
            K.append(np.broadcast_to(k, self.shape(True)) if broadcast else k)
        return K

    def get_mask_nyquist(self):
        mask = np.ones(self.global_shape(True), dtype=int)
        for i, n in enumerate(self.N):
            if n % 2 == 0:
                s = [slice(None)]*self.dim
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
        CompositeSpace.__init__(self, [T]*len(T.N))


def _scalar_space(space):
    return space.T if isinstance(space, CompositeSpace) else space


# ---- lazy host mirror of the device-resident state (opt-in: SDNS_LAZY_STATE=1) -------------------------------------
# While solve() runs, the solution lives on the GPU and the context's numpy array is a mirror.  Eagerly, solve()
# refreshes the mirror before every user callback (one D2H per step) and assumes the callback wrote it (one H2D).  In
# lazy mode the copy is deferred until host code touches the array through numpy's protocols, and the expressions
# demo/Isotropic.py's update() uses every step are answered on the device (spectraldns_b200/diagnostics.py), so a
# forced run moves no state across PCIe.  What numpy does not route through a Python hook (the buffer protocol:
# memoryview, Cython typed views, views taken before solve()) still sees the mirror as of the last refresh -- hence
# opt-in.
_LAZY = []          # DeviceStates whose host mirror may be stale


def lazy_state_enabled():
    import os
    return os.environ.get('SDNS_LAZY_STATE', '0') not in ('', '0')


def _lazy_dev(a):
    """The DeviceState whose registered host array shares memory with `a`, if its mirror is lazily maintained."""
    if not _LAZY or not isinstance(a, np.ndarray):
        return None
    try:
        lo = a.__array_interface__['data'][0]
    except Exception:
        return None
    for dev in _LAZY:
        h = dev.host_state
        b0 = h.__array_interface__['data'][0]
        if b0 <= lo < b0 + h.nbytes:
            return dev
    return None


def _is_whole_state(a, dev):
    h = dev.host_state
    return (a.shape == h.shape and a.dtype == h.dtype and a.flags['C_CONTIGUOUS']
            and a.__array_interface__['data'][0] == h.__array_interface__['data'][0])


class _ScaledState(object):
    """U_hat*weight, not formed: energy_fourier() of it runs on the device; anything else materialises it."""
    __array_priority__ = 100

    def __init__(self, dev, state, weight):
        self.dev, self.state, self.weight = dev, state, weight

    def materialise(self):
        self.dev.sync_to_host()
        return np.asarray(self.state.view(np.ndarray))*self.weight

    def __array__(self, dtype=None, copy=None):
        a = self.materialise()
        return a.astype(dtype) if dtype is not None else a

    def __getattr__(self, name):            # any ndarray attribute / method: act as the product
        return getattr(self.materialise(), name)

    def __mul__(self, o):
        return self.materialise()*o
    __rmul__ = __mul__

    def __getitem__(self, k):
        return self.materialise()[k]


class _SpaceArray(np.ndarray):
    """ndarray that remembers its function space; slices and views keep it (UB[:3], U_hat[0])."""
    _spectral = False

    # -- lazy mirror hooks (inactive unless a DeviceState registered itself in _LAZY) --
    def _lazy_read(self):
        dev = _lazy_dev(self)
        if dev is not None:
            dev.sync_to_host()
        return dev

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        if _LAZY:
            outs = out if isinstance(out, tuple) else ((out,) if out is not None else ())
            wdev = None
            for o in outs:
                d = _lazy_dev(o) if isinstance(o, np.ndarray) else None
                if d is not None:
                    wdev = d
            # U_hat *= real field  (demo/Isotropic.py:180) and U_hat*real field (Isotropic.py:168)
            if ufunc is np.multiply and method == '__call__' and len(inputs) == 2 and isinstance(inputs[0], _SpaceArray):
                dev = _lazy_dev(inputs[0])
                w = inputs[1]
                if dev is not None and _is_whole_state(inputs[0], dev) and isinstance(w, np.ndarray) and not isinstance(w, _SpaceArray) \
                        and w.dtype.kind in 'fiub' and w.shape == tuple(dev.plan.spectral_shape):
                    from . import diagnostics
                    if len(outs) == 1 and outs[0] is inputs[0]:
                        if dev.host_dirty:
                            dev.upload_state()
                        dev.plan.scale_field(dev.u, diagnostics.real_field(dev, w, 'imul'))
                        dev.device_newer = True
                        return inputs[0]
                    if not outs:
                        return _ScaledState(dev, inputs[0], w)
            for a in inputs:
                if isinstance(a, np.ndarray):
                    d = _lazy_dev(a)
                    if d is not None:
                        d.sync_to_host()
            if wdev is not None:
                wdev.sync_to_host()
                wdev.host_touched()
        # ndarray's own implementation on plain views, results re-wrapped (the pattern of the numpy subclassing guide)
        first = next((a for a in inputs if isinstance(a, _SpaceArray)), None)
        args = [a.view(np.ndarray) if isinstance(a, _SpaceArray) else a for a in inputs]
        outs_in = out if isinstance(out, tuple) else ((out,) if out is not None else None)
        if outs_in is not None:
            kw['out'] = tuple(o.view(np.ndarray) if isinstance(o, _SpaceArray) else o for o in outs_in)
        res = super().__array_ufunc__(ufunc, method, *args, **kw)
        if res is NotImplemented or method == 'at':
            return res
        many = isinstance(res, tuple)
        rs = list(res) if many else [res]
        for i, r in enumerate(rs):
            if outs_in is not None and i < len(outs_in) and outs_in[i] is not None:
                rs[i] = outs_in[i]
            elif isinstance(r, np.ndarray) and first is not None:
                w = r.view(type(first))
                w._space = first._space
                rs[i] = w
        return tuple(rs) if many else rs[0]

    def __array_function__(self, func, types, args, kwargs):
        if _LAZY:
            for a in list(args) + list(kwargs.values()):
                if isinstance(a, np.ndarray):
                    d = _lazy_dev(a)
                    if d is not None:
                        d.sync_to_host()
                        d.host_touched()            # conservative: the function may write through the array
        return super().__array_function__(func, types, args, kwargs)

    def __getitem__(self, key):
        if _LAZY:
            self._lazy_read()
        return super().__getitem__(key)

    def __setitem__(self, key, value):
        if _LAZY:
            dev = _lazy_dev(self)
            if dev is not None:
                # U_hat[:, i0, i1, i2] = scalar (demo/Isotropic.py:64, 163: the mean mode)
                if _is_whole_state(self, dev) and isinstance(key, tuple) and len(key) == 4 and isinstance(key[0], slice) \
                        and key[0] == slice(None) \
                        and all(isinstance(k, (int, np.integer)) for k in key[1:]) and np.isscalar(value):
                    if dev.host_dirty:
                        dev.upload_state()
                    s = dev.plan.spectral_shape
                    idx = tuple(int(k) % n for k, n in zip(key[1:], s))
                    dev.plan.set_mode(dev.u, idx, value)
                    self.view(np.ndarray)[key] = value       # keep the mirror's entry in step (plain view: no hooks)
                    return
                dev.sync_to_host()
                dev.host_touched()
                self.view(np.ndarray)[key] = value
                return
        super().__setitem__(key, value)

    def copy(self, *a, **k):
        if _LAZY:
            self._lazy_read()
        return super().copy(*a, **k)

    def astype(self, *a, **k):
        if _LAZY:
            self._lazy_read()
        return super().astype(*a, **k)

    def tobytes(self, *a, **k):
        if _LAZY:
            self._lazy_read()
        return super().tobytes(*a, **k)

    def fill(self, v):
        if _LAZY:
            dev = _lazy_dev(self)
            if dev is not None:
                dev.sync_to_host()
                dev.host_touched()
        return super().fill(v)

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
        return T if self.ndim == len(T.N) else CompositeSpace([T]*self.shape[0])


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
    if _LAZY:
        # the device-resident state itself, or state*weight (demo/Isotropic.py:167-168): reduce it where it lives
        st, w = (u_hat.state, u_hat.weight) if isinstance(u_hat, _ScaledState) else (u_hat, None)
        dev = _lazy_dev(st) if isinstance(st, np.ndarray) else None
        if dev is not None and _is_whole_state(st, dev):
            from . import diagnostics
            if dev.host_dirty:
                dev.upload_state()
            dev.plan.use_current_stream()
            res = dev.plan.energy_weighted(dev.u, diagnostics.real_field(dev, w, 'energy_weight') if w is not None else None)
            comm = getattr(S, 'comm', None)
            if comm is not None and hasattr(comm, 'allreduce'):
                res = comm.allreduce(res)
            return float(res)
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
