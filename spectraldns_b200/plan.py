"""Plan: the Python face of one sdns_plan (include/sdns_b200.h).

PyTorch is used only to allocate device buffers and to name the CUDA stream; every arithmetic
operation goes through the C ABI into the hand-written sm_100a kernels.  There is no CPU path.
"""
import ctypes as C
import os
import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


class Plan(object):
    """One grid / precision / solver configuration bound to one GPU.

    Mirrors what the reference's get_context() builds (solvers/NS.py:12-48): the spaces T and
    Tp, wavenumbers, masks -- all of which live inside the C plan -- plus scratch memory.
    """

    def __init__(self, N, L=(2*np.pi,)*3, precision='double', dealias='2/3-rule', solver='NS',
                 convection=None, mask_nyquist=True, decomposition='slab', kcut=None, device=0,
                 rank=0, nranks=1, k1_layout=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.SdnsError('no CUDA device: spectraldns_b200 has no CPU fallback')
        self.lib = _lib.lib()
        self.N = tuple(int(n) for n in N)
        self.L = tuple(float(l) for l in L)
        self.precision = precision
        self.dealias = dealias
        self.solver = solver
        if convection is None:
            convection = 'Divergence' if solver == 'MHD' else 'Vortex'
        self.convection = convection
        self.float = np.dtype(np.float32 if precision == 'single' else np.float64)
        self.complex = np.dtype(np.complex64 if precision == 'single' else np.complex128)
        self.tfloat = torch.float32 if precision == 'single' else torch.float64
        self.tcomplex = torch.complex64 if precision == 'single' else torch.complex128
        self.device = torch.device('cuda', device)
        cfg = _lib.SdnsConfig()
        cfg.abi_version = _lib.SDNS_ABI_VERSION
        for i in range(3):
            cfg.N[i] = self.N[i]
            cfg.L[i] = self.L[i]
            cfg.kcut[i] = -1 if kcut is None else int(kcut[i])
        cfg.precision = _lib.SINGLE if precision == 'single' else _lib.DOUBLE
        cfg.dealias = _lib.DEALIAS[dealias]
        cfg.solver = _lib.SOLVER[solver]
        cfg.convection = _lib.CONVECTION[convection]
        cfg.mask_nyquist = 1 if mask_nyquist else 0
        cfg.decomposition = _lib.DECOMP[decomposition]
        cfg.prune = 1
        cfg.rank, cfg.nranks, cfg.device = int(rank), int(nranks), device
        self.rank, self.nranks = int(rank), int(nranks)
        # which axis-1 modes a rank owns: 'blocks' = the reference's contiguous slabs (spectralinit.py:19-21),
        # 'cyclic' = rank r owns [r::P] (balances the modes the 2/3 rule keeps; include/sdns_b200.h)
        if k1_layout is None:
            k1_layout = os.environ.get('SDNS_K1_LAYOUT', 'blocks')
        cfg.k1_layout = _lib.K1_LAYOUT[k1_layout]
        self.k1_layout = k1_layout if self.nranks > 1 else 'blocks'
        self._p = C.c_void_p()
        _lib.check(self.lib.sdns_plan_create(C.byref(self._p), C.byref(cfg)))
        sp, ph, pd = (C.c_int32*3)(), (C.c_int32*3)(), (C.c_int32*3)()
        _lib.check(self.lib.sdns_local_shapes(self._p, C.byref(sp), C.byref(ph), C.byref(pd)))
        k1a, k1s = C.c_int32(), C.c_int32()
        _lib.check(self.lib.sdns_k1_layout(self._p, C.byref(k1a), C.byref(k1s)))
        # T.local_slice(True)[1]: the axis-1 modes of this rank as a slice of the global axis
        self.k1_slice = (slice(k1a.value, self.N[1], k1s.value) if k1s.value > 1 else slice(k1a.value, k1a.value + sp[1]))
        self.spectral_shape = tuple(sp)
        self.physical_shape = tuple(ph)
        self.padded_shape = tuple(pd)
        # this rank's planes of the physical axis 0, plain and dealiased space (M0 // P per rank, the first M0 % P one more)
        def share(m):
            c, rem = divmod(m, self.nranks)
            lo = self.rank*c + min(self.rank, rem)
            return slice(lo, lo + c + (1 if self.rank < rem else 0))
        M0p = self.N[0]*3//2 if dealias == '3/2-rule' else self.N[0]
        self.x0_slice, self.x0p_slice = share(self.N[0]), share(M0p)
        self.ncomp = 6 if solver == 'MHD' else 3
        nb = C.c_size_t()
        _lib.check(self.lib.sdns_workspace_bytes(self._p, C.byref(nb)))
        self.workspace_bytes = nb.value
        with torch.cuda.device(self.device):
            self.use_current_stream()
            if self.nranks == 1:
                # PyTorch is the device-memory allocator
                self._ws = torch.empty(nb.value + 256, dtype=torch.uint8, device=self.device)
                off = (-self._ws.data_ptr()) % 256
                _lib.check(self.lib.sdns_plan_set_workspace(self._p, self._ws.data_ptr() + off, nb.value))
            else:
                self._open_peers()

    def _open_peers(self):
        """Slab decomposition: the library cudaMallocs the workspace so that it can be shared with
        the other ranks through CUDA IPC; handles travel over torch.distributed."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.SdnsError('nranks > 1 needs an initialised torch.distributed process group')
        if dist.get_world_size() != self.nranks or dist.get_rank() != self.rank:
            raise _lib.SdnsError('rank/nranks do not match the torch.distributed process group')
        _lib.check(self.lib.sdns_comm_alloc(self._p))
        buf = (C.c_char*64)()
        _lib.check(self.lib.sdns_comm_handle(self._p, buf))
        mine = bytes(buf.raw)
        handles = [None]*self.nranks
        dist.all_gather_object(handles, mine)
        blob = b''.join(handles)
        _lib.check(self.lib.sdns_comm_open(self._p, C.c_char_p(blob), self.nranks))
        dist.barrier()

    def comm_timed_out(self):
        v = C.c_int()
        _lib.check(self.lib.sdns_comm_status(self._p, C.byref(v)))
        return bool(v.value)

    def __del__(self):
        try:
            if getattr(self, '_p', None):
                self.lib.sdns_plan_destroy(self._p)
                self._p = None
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------
    def use_current_stream(self):
        """Enqueue on torch's current stream (so torch.cuda.Event brackets the kernels)."""
        torch = _torch()
        s = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.sdns_plan_set_stream(self._p, C.c_void_p(s)))

    def sync(self):
        _lib.check(self.lib.sdns_sync(self._p))

    def launch_count(self):
        c = C.c_longlong()
        _lib.check(self.lib.sdns_launch_count(self._p, C.byref(c)))
        return c.value

    def empty_spectral(self, ncomp=None):
        torch = _torch()
        shape = self.spectral_shape if ncomp == 0 else ((ncomp or self.ncomp),) + self.spectral_shape
        return torch.zeros(shape, dtype=self.tcomplex, device=self.device)

    def empty_physical(self, ncomp=None, padded=False):
        torch = _torch()
        s = self.padded_shape if padded else self.physical_shape
        shape = s if ncomp == 0 else ((ncomp or self.ncomp),) + s
        return torch.zeros(shape, dtype=self.tfloat, device=self.device)

    def to_device(self, a, complex_=None):
        torch = _torch()
        a = np.ascontiguousarray(a)
        if np.iscomplexobj(a):
            a = a.astype(self.complex, copy=False)
        else:
            a = a.astype(self.float, copy=False)
        return torch.from_numpy(a).to(self.device)

    @staticmethod
    def to_host(t):
        return t.detach().cpu().numpy()

    def _chk(self, t, dtype, shape_tail, name):
        if t.dtype != dtype or not t.is_contiguous() or not t.is_cuda:
            raise ValueError('%s: need a contiguous CUDA tensor of dtype %s' % (name, dtype))
        if tuple(t.shape[-3:]) != tuple(shape_tail):
            raise ValueError('%s: trailing shape %s != %s' % (name, tuple(t.shape[-3:]), tuple(shape_tail)))
        return t.numel() // int(np.prod(shape_tail))

    # -- transforms (T/VT and Tp/VTp forward/backward) ---------------------
    def forward(self, u, out=None, padded=False):
        s = self.padded_shape if padded else self.physical_shape
        nc = self._chk(u, self.tfloat, s, 'forward input')
        if out is None:
            out = _torch().empty(tuple(u.shape[:-3]) + self.spectral_shape, dtype=self.tcomplex, device=self.device)
        assert self._chk(out, self.tcomplex, self.spectral_shape, 'forward output') == nc
        _lib.check(self.lib.sdns_forward(self._p, _lib.SPACE_TP if padded else _lib.SPACE_T, nc,
                                         u.data_ptr(), out.data_ptr()))
        return out

    def backward(self, u_hat, out=None, padded=False, dealias=False):
        """padded/dealias select the reference's Tp space (solvers/NS.py:29-32)."""
        use_tp = padded or dealias
        s = self.padded_shape if use_tp else self.physical_shape
        nc = self._chk(u_hat, self.tcomplex, self.spectral_shape, 'backward input')
        if out is None:
            out = _torch().empty(tuple(u_hat.shape[:-3]) + s, dtype=self.tfloat, device=self.device)
        assert self._chk(out, self.tfloat, s, 'backward output') == nc
        _lib.check(self.lib.sdns_backward(self._p, _lib.SPACE_TP if use_tp else _lib.SPACE_T, nc,
                                          u_hat.data_ptr(), out.data_ptr()))
        return out

    # -- the hot path --------------------------------------------------------
    def compute_rhs(self, rhs, u_hat, nu, eta=0.0, source=None, p_hat=None):
        assert self._chk(rhs, self.tcomplex, self.spectral_shape, 'rhs') == self.ncomp
        assert self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat') == self.ncomp
        _lib.check(self.lib.sdns_compute_rhs(self._p, rhs.data_ptr(), u_hat.data_ptr(), float(nu), float(eta),
                                             source.data_ptr() if source is not None else None,
                                             p_hat.data_ptr() if p_hat is not None else None))
        return rhs

    def compute_conv(self, rhs, u_hat):
        """solver.conv: the dealiased nonlinear term only."""
        _lib.check(self.lib.sdns_compute_conv(self._p, rhs.data_ptr(), u_hat.data_ptr()))
        return rhs

    def rk4_step(self, u_hat, u1, u2, dt, nu, eta=0.0, source=None):
        for t, n in ((u_hat, 'u_hat'), (u1, 'u1'), (u2, 'u2')):
            assert self._chk(t, self.tcomplex, self.spectral_shape, n) == self.ncomp
        _lib.check(self.lib.sdns_rk4_step(self._p, u_hat.data_ptr(), u1.data_ptr(), u2.data_ptr(), float(dt),
                                          float(nu), float(eta),
                                          source.data_ptr() if source is not None else None))
        return u_hat

    def euler_step(self, u_hat, rhs, dt, nu, eta=0.0, source=None):
        _lib.check(self.lib.sdns_euler_step(self._p, u_hat.data_ptr(), rhs.data_ptr(), float(dt), float(nu),
                                            float(eta), source.data_ptr() if source is not None else None))
        return u_hat

    def ab2_step(self, u_hat, u1, rhs, dt, tstep, nu, eta=0.0, source=None):
        _lib.check(self.lib.sdns_ab2_step(self._p, u_hat.data_ptr(), u1.data_ptr(), rhs.data_ptr(), float(dt),
                                          int(tstep), float(nu), float(eta),
                                          source.data_ptr() if source is not None else None))
        return u_hat

    def rk4_steps_host(self, host_u_hat, u_hat, u1, u2, nsteps, dt, nu, eta=0.0):
        """integrate() for a caller that keeps the state in (pinned) host memory: H2D, nsteps RK4
        steps, D2H.  host_u_hat is a torch CPU tensor (ideally pinned) or a numpy array."""
        ptr = host_u_hat.data_ptr() if hasattr(host_u_hat, 'data_ptr') else host_u_hat.ctypes.data
        _lib.check(self.lib.sdns_rk4_steps_host(self._p, ptr, u_hat.data_ptr(), u1.data_ptr(), u2.data_ptr(),
                                                int(nsteps), float(dt), float(nu), float(eta)))
        return host_u_hat

    def cross2(self, c, b, over_k2=False):
        _lib.check(self.lib.sdns_cross2(self._p, c.data_ptr(), b.data_ptr(), 1 if over_k2 else 0))
        return c

    def cross1(self, c, a, b):
        _lib.check(self.lib.sdns_cross1(self._p, c.data_ptr(), a.data_ptr(), b.data_ptr(), a.numel()//3))
        return c

    def cross2_dense(self, c, a, b):
        _lib.check(self.lib.sdns_cross2_dense(self._p, c.data_ptr(), a.data_ptr(), b.data_ptr()))
        return c

    def project(self, u_hat):
        _lib.check(self.lib.sdns_project(self._p, u_hat.data_ptr()))
        return u_hat

    def add_pressure_diffusion(self, du, u_hat, nu, p_hat=None):
        """add_pressure_diffusion_NS (solvers/NS.py:203-217, cython_solvers.in:40-80) on its own: in place on du."""
        for t, n in ((du, 'du'), (u_hat, 'u_hat')):
            assert self._chk(t, self.tcomplex, self.spectral_shape, n) == 3
        _lib.check(self.lib.sdns_add_pressure_diffusion(self._p, du.data_ptr(), u_hat.data_ptr(), float(nu),
                                                        p_hat.data_ptr() if p_hat is not None else None))
        return du

    def lincomb(self, out, base, coeffs, arrays):
        """out = base + sum_t coeffs[t]*arrays[t] (base may be None); <= 9 terms."""
        n = len(coeffs)
        cs = (C.c_double*max(n, 1))(*[float(c) for c in coeffs])
        ps = (C.c_void_p*max(n, 1))(*[a.data_ptr() for a in arrays])
        nc = out.numel() // int(np.prod(self.spectral_shape))
        _lib.check(self.lib.sdns_lincomb(self._p, out.data_ptr(), base.data_ptr() if base is not None else None,
                                         n, cs, ps, nc))
        return out

    def errnorm(self, u0, u1, err, atol, rtol):
        nc = u0.numel() // int(np.prod(self.spectral_shape))
        out = (C.c_double*nc)()
        _lib.check(self.lib.sdns_errnorm(self._p, u0.data_ptr(), u1.data_ptr(), err.data_ptr(), float(atol),
                                         float(rtol), nc, out))
        return np.array(out[:])

    def energy(self, u_hat):
        nc = self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat')
        out = C.c_double()
        _lib.check(self.lib.sdns_energy(self._p, u_hat.data_ptr(), nc, C.byref(out)))
        return out.value

    # -- diagnostics / forcing of demo/Isotropic.py on the device -----------------
    def _real_field(self, w, name):
        torch = _torch()
        if w.dtype not in (torch.float32, torch.float64) or not w.is_cuda or not w.is_contiguous() \
                or tuple(w.shape) != tuple(self.spectral_shape):
            raise ValueError('%s: need a contiguous real CUDA tensor of the spectral shape %s' % (name, self.spectral_shape))
        return 1 if w.dtype == torch.float64 else 0

    def energy_weighted(self, u_hat, weight=None):
        """energy_fourier(u_hat*weight, T) of the local block (demo/Isotropic.py:168)."""
        nc = self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat')
        out = C.c_double()
        isd = self._real_field(weight, 'weight') if weight is not None else 1
        _lib.check(self.lib.sdns_energy_weighted(self._p, u_hat.data_ptr(), nc, weight.data_ptr() if weight is not None else None,
                                                 isd, C.byref(out)))
        return out.value

    def scale_field(self, u_hat, factor, a=1.0, b=0.0):
        """u_hat *= a*factor + b*(1 - factor), factor a real field broadcast over the components: the plain product by
        default, the forcing rescale of demo/Isotropic.py:180 with factor = k2_mask, (a, b) = (alpha, 1)."""
        nc = self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat')
        _lib.check(self.lib.sdns_scale_field(self._p, u_hat.data_ptr(), nc, factor.data_ptr(), self._real_field(factor, 'factor'),
                                             float(a), float(b)))
        return u_hat

    def set_mode(self, u_hat, index, value=0.0):
        """u_hat[:, i0, i1, i2] = value (i1 local)."""
        nc = self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat')
        v = complex(value)
        _lib.check(self.lib.sdns_set_mode(self._p, u_hat.data_ptr(), nc, int(index[0]), int(index[1]), int(index[2]), v.real, v.imag))
        return u_hat

    def enstrophy(self, u_hat):
        """energy_fourier(1j*K x u_hat, T) of the local block: the `dissipation` of demo/Isotropic.py:243-244."""
        assert self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat') == 3
        out = C.c_double()
        _lib.check(self.lib.sdns_enstrophy(self._p, u_hat.data_ptr(), C.byref(out)))
        return out.value

    def divergence_norm(self, u_hat):
        """sum w |1j*K.u_hat|^2 of the local block (= L2_norm(get_divergence) of demo/Isotropic.py:245-247 by Parseval)."""
        assert self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat') == 3
        out = C.c_double()
        _lib.check(self.lib.sdns_divergence_norm(self._p, u_hat.data_ptr(), C.byref(out)))
        return out.value

    def spectrum_shells(self, u_hat, nbins):
        """(sums, counts) over the shells of spectrum() (demo/Isotropic.py:88-118), local block."""
        nc = self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat')
        s, c = (C.c_double*nbins)(), (C.c_double*nbins)()
        _lib.check(self.lib.sdns_spectrum(self._p, u_hat.data_ptr(), nc, int(nbins), s, c))
        return np.array(s[:]), np.array(c[:])

    # -- measurement ---------------------------------------------------------
    FAMILIES = ['plain_fwd_c2c', 'plain_bwd_c2c', 'ns_b0', 'vv_b0', 'ns_f0', 'vv_f0', 'mhd_f0',
                'c2r', 'r2c', 'z_cross', 'z_mhd', 'ns_grad_b0', 'z_dot', 'z_uu', 'nsdiv_f0']

    def profile(self, on=True, timeline=False):
        _lib.check(self.lib.sdns_profile_enable(self._p, (2 if timeline else 1) if on else 0))

    def profile_timeline(self):
        """[(kind, t_start_ms, t_end_ms, bytes)] since profile(True, timeline=True); kind is a FAMILIES name,
        'barrier' or 'copy<stream>'."""
        n = C.c_int()
        _lib.check(self.lib.sdns_profile_timeline(self._p, None, 0, C.byref(n)))
        # the first call consumed the pending records into the timeline; fetch it
        buf = (C.c_double*(4*max(n.value, 1)))()
        m = C.c_int()
        _lib.check(self.lib.sdns_profile_timeline(self._p, buf, n.value, C.byref(m)))
        out = []
        for i in range(min(n.value, m.value)):
            k = int(buf[4*i])
            name = self.FAMILIES[k] if 0 <= k < len(self.FAMILIES) else ('barrier' if k == 99 else ('xfer' if k == 98 else 'copy%d' % (k-100)))
            out.append((name, buf[4*i+1], buf[4*i+2], buf[4*i+3]))
        return out

    def profile_read(self):
        """{family: (total_ms, launches, algorithmic_hbm_bytes, nvlink_bytes_stored)} since profile(True)."""
        out = {}
        for i, name in enumerate(self.FAMILIES):
            ms, n, b = C.c_double(), C.c_longlong(), C.c_double()
            _lib.check(self.lib.sdns_profile_read(self._p, i, C.byref(ms), C.byref(n), C.byref(b)))
            if n.value:
                r = C.c_double()
                _lib.check(self.lib.sdns_profile_read_nvlink(self._p, i, C.byref(r)))
                out[name] = (ms.value, n.value, b.value, r.value)
        return out

    def xfer_stats(self):
        """(bytes this rank has sent over NVLink through the transfer role since plan creation, transfer-only launches)."""
        b, n = C.c_double(), C.c_longlong()
        _lib.check(self.lib.sdns_xfer_stats(self._p, C.byref(b), C.byref(n)))
        return b.value, n.value

    def profile_read_copies(self):
        """(busy_ms of the busiest per-peer copy stream, bytes sent over NVLink, number of copies) since
        profile(True); all zero unless the copy engines carry the exchange (nranks > 1, the default)."""
        ms, b, n = C.c_double(), C.c_double(), C.c_longlong()
        _lib.check(self.lib.sdns_profile_read_copies(self._p, C.byref(ms), C.byref(b), C.byref(n)))
        return ms.value, b.value, n.value


class Plan2D(object):
    """One doubly periodic grid / precision / solver (NS2D or Bq2D) bound to one GPU: the Python face of sdns2d_plan.

    Mirrors what get_context() of solvers/NS2D.py:13-18 and Bq2D.py:13-50 builds.  Method names follow Plan, so that
    the host layer (device_state.py, compat/spectralDNS/maths) drives both alike; `eta` arguments are ignored and the
    Boussinesq numbers Ri, Pr are attributes (set_physics)."""

    def __init__(self, N, L=(2*np.pi,)*2, precision='double', dealias='2/3-rule', solver='NS2D', mask_nyquist=True,
                 kcut=None, device=0):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.SdnsError('no CUDA device: spectraldns_b200 has no CPU fallback')
        self.lib = _lib.lib()
        self.N = tuple(int(n) for n in N)
        self.L = tuple(float(l) for l in L)
        self.precision, self.dealias, self.solver = precision, dealias, solver
        self.convection = 'Vortex'
        self.float = np.dtype(np.float32 if precision == 'single' else np.float64)
        self.complex = np.dtype(np.complex64 if precision == 'single' else np.complex128)
        self.tfloat = torch.float32 if precision == 'single' else torch.float64
        self.tcomplex = torch.complex64 if precision == 'single' else torch.complex128
        self.device = torch.device('cuda', device)
        self.rank, self.nranks = 0, 1
        self.Ri, self.Pr = 0.1, 1.0                      # config.py:260-261 defaults
        cfg = _lib.Sdns2dConfig()
        cfg.abi_version = _lib.SDNS_ABI_VERSION
        for i in range(2):
            cfg.N[i], cfg.L[i] = self.N[i], self.L[i]
            cfg.kcut[i] = -1 if kcut is None else int(kcut[i])
        cfg.precision = _lib.SINGLE if precision == 'single' else _lib.DOUBLE
        cfg.dealias = _lib.DEALIAS[dealias]
        cfg.solver = _lib.SOLVER2D[solver]
        cfg.mask_nyquist = 1 if mask_nyquist else 0
        cfg.device = device
        self._p = C.c_void_p()
        _lib.check2d(self.lib.sdns2d_plan_create(C.byref(self._p), C.byref(cfg)))
        sp, ph, pd = (C.c_int32*2)(), (C.c_int32*2)(), (C.c_int32*2)()
        _lib.check2d(self.lib.sdns2d_shapes(self._p, C.byref(sp), C.byref(ph), C.byref(pd)))
        self.spectral_shape, self.physical_shape, self.padded_shape = tuple(sp), tuple(ph), tuple(pd)
        self.ncomp = 3 if solver == 'Bq2D' else 2
        nb = C.c_size_t()
        _lib.check2d(self.lib.sdns2d_workspace_bytes(self._p, C.byref(nb)))
        self.workspace_bytes = nb.value
        with torch.cuda.device(self.device):
            self.use_current_stream()
            self._ws = torch.empty(nb.value + 256, dtype=torch.uint8, device=self.device)
            off = (-self._ws.data_ptr()) % 256
            _lib.check2d(self.lib.sdns2d_plan_set_workspace(self._p, self._ws.data_ptr() + off, nb.value))

    def __del__(self):
        try:
            if getattr(self, '_p', None):
                self.lib.sdns2d_plan_destroy(self._p)
                self._p = None
        except Exception:
            pass

    def set_physics(self, params):
        """Richardson and Prandtl numbers of the Boussinesq solver from the run-time parameters (config.py:260-261)."""
        if 'Ri' in params:
            self.Ri = float(params.Ri)
        if 'Pr' in params:
            self.Pr = float(params.Pr)

    def use_current_stream(self):
        s = _torch().cuda.current_stream(self.device).cuda_stream
        _lib.check2d(self.lib.sdns2d_plan_set_stream(self._p, C.c_void_p(s)))

    def sync(self):
        _lib.check2d(self.lib.sdns2d_sync(self._p))

    def comm_timed_out(self):
        return False

    def launch_count(self):
        c = C.c_longlong()
        _lib.check2d(self.lib.sdns2d_launch_count(self._p, C.byref(c)))
        return c.value

    def empty_spectral(self, ncomp=None):
        shape = self.spectral_shape if ncomp == 0 else ((ncomp or self.ncomp),) + self.spectral_shape
        return _torch().zeros(shape, dtype=self.tcomplex, device=self.device)

    def empty_physical(self, ncomp=None, padded=False):
        s = self.padded_shape if padded else self.physical_shape
        shape = s if ncomp == 0 else ((ncomp or self.ncomp),) + s
        return _torch().zeros(shape, dtype=self.tfloat, device=self.device)

    to_device = Plan.to_device
    to_host = staticmethod(Plan.to_host)

    def _chk(self, t, dtype, shape_tail, name):
        if t.dtype != dtype or not t.is_contiguous() or not t.is_cuda:
            raise ValueError('%s: need a contiguous CUDA tensor of dtype %s' % (name, dtype))
        if tuple(t.shape[-2:]) != tuple(shape_tail):
            raise ValueError('%s: trailing shape %s != %s' % (name, tuple(t.shape[-2:]), tuple(shape_tail)))
        return t.numel() // int(np.prod(shape_tail))

    def forward(self, u, out=None, padded=False):
        s = self.padded_shape if padded else self.physical_shape
        nc = self._chk(u, self.tfloat, s, 'forward input')
        if out is None:
            out = _torch().empty(tuple(u.shape[:-2]) + self.spectral_shape, dtype=self.tcomplex, device=self.device)
        assert self._chk(out, self.tcomplex, self.spectral_shape, 'forward output') == nc
        _lib.check2d(self.lib.sdns2d_forward(self._p, _lib.SPACE_TP if padded else _lib.SPACE_T, nc, u.data_ptr(), out.data_ptr()))
        return out

    def backward(self, u_hat, out=None, padded=False, dealias=False):
        use_tp = padded or dealias
        s = self.padded_shape if use_tp else self.physical_shape
        nc = self._chk(u_hat, self.tcomplex, self.spectral_shape, 'backward input')
        if out is None:
            out = _torch().empty(tuple(u_hat.shape[:-2]) + s, dtype=self.tfloat, device=self.device)
        assert self._chk(out, self.tfloat, s, 'backward output') == nc
        _lib.check2d(self.lib.sdns2d_backward(self._p, _lib.SPACE_TP if use_tp else _lib.SPACE_T, nc, u_hat.data_ptr(), out.data_ptr()))
        return out

    def compute_rhs(self, rhs, u_hat, nu, eta=0.0, source=None, p_hat=None):
        assert self._chk(rhs, self.tcomplex, self.spectral_shape, 'rhs') == self.ncomp
        assert self._chk(u_hat, self.tcomplex, self.spectral_shape, 'u_hat') == self.ncomp
        _lib.check2d(self.lib.sdns2d_compute_rhs(self._p, rhs.data_ptr(), u_hat.data_ptr(), float(nu), self.Ri, self.Pr,
                                                 source.data_ptr() if source is not None else None,
                                                 p_hat.data_ptr() if p_hat is not None else None))
        return rhs

    def rk4_step(self, u_hat, u1, u2, dt, nu, eta=0.0, source=None):
        for t, n in ((u_hat, 'u_hat'), (u1, 'u1'), (u2, 'u2')):
            assert self._chk(t, self.tcomplex, self.spectral_shape, n) == self.ncomp
        _lib.check2d(self.lib.sdns2d_rk4_step(self._p, u_hat.data_ptr(), u1.data_ptr(), u2.data_ptr(), float(dt), float(nu),
                                              self.Ri, self.Pr, source.data_ptr() if source is not None else None))
        return u_hat

    def euler_step(self, u_hat, rhs, dt, nu, eta=0.0, source=None):
        _lib.check2d(self.lib.sdns2d_euler_step(self._p, u_hat.data_ptr(), rhs.data_ptr(), float(dt), float(nu), self.Ri, self.Pr,
                                                source.data_ptr() if source is not None else None))
        return u_hat

    def ab2_step(self, u_hat, u1, rhs, dt, tstep, nu, eta=0.0, source=None):
        _lib.check2d(self.lib.sdns2d_ab2_step(self._p, u_hat.data_ptr(), u1.data_ptr(), rhs.data_ptr(), float(dt), int(tstep),
                                              float(nu), self.Ri, self.Pr, source.data_ptr() if source is not None else None))
        return u_hat

    def cross2(self, c, u_hat):
        """Scalar c = 1j*(K0 u1 - K1 u0) (solvers/NS2D.py:20-23)."""
        _lib.check2d(self.lib.sdns2d_cross2(self._p, c.data_ptr(), u_hat.data_ptr()))
        return c

    def add_pressure_diffusion(self, du, u_hat, nu, p_hat=None):
        _lib.check2d(self.lib.sdns2d_add_pressure_diffusion(self._p, du.data_ptr(), u_hat.data_ptr(), float(nu), self.Ri, self.Pr,
                                                            p_hat.data_ptr() if p_hat is not None else None))
        return du
