"""TEST INFRASTRUCTURE ONLY -- numpy front end of tests/host/build/libsdns_emu.so (the library's own sources compiled
against the CUDA-model emulation in host_shim.h).  "Device" pointers are numpy buffers.  Mirrors the calls of
spectraldns_b200/plan.py that the parity tests need."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from spectraldns_b200 import _lib   # noqa: E402  (SdnsConfig / enums only; the emulated library is loaded below)

vp = C.c_void_p


def load(path):
    L = C.CDLL(path)
    L.sdns_last_error.restype = C.c_char_p
    L.sdns_plan_create.argtypes = [C.POINTER(vp), C.POINTER(_lib.SdnsConfig)]
    L.sdns_plan_destroy.argtypes = [vp]
    L.sdns_workspace_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.sdns_plan_set_workspace.argtypes = [vp, vp, C.c_size_t]
    L.sdns_forward.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.sdns_backward.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.sdns_compute_rhs.argtypes = [vp, vp, vp, C.c_double, C.c_double, vp, vp]
    L.sdns_compute_conv.argtypes = [vp, vp, vp]
    L.sdns_rk4_step.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, vp]
    L.sdns_energy.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_double)]
    L.sdns_comm_alloc.argtypes = [vp]
    L.sdns_comm_handle.argtypes = [vp, vp]
    L.sdns_comm_open.argtypes = [vp, vp, C.c_int]
    L.sdns_comm_status.argtypes = [vp, C.POINTER(C.c_int)]
    if hasattr(L, 'sdns_emu_set_skew'):
        L.sdns_emu_set_skew.argtypes = [C.c_int]
        L.sdns_emu_set_skew.restype = None
    L.sdns_local_shapes.argtypes = [vp, C.POINTER(C.c_int32*3), C.POINTER(C.c_int32*3), C.POINTER(C.c_int32*3)]
    L.sdns_k1_layout.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    return L


class EmuPlan(object):
    def __init__(self, L, N, Lbox=(2*np.pi,)*3, precision='double', dealias='2/3-rule', solver='NS',
                 convection=None, mask_nyquist=True, kcut=None, rank=0, nranks=1, k1_layout='blocks'):
        self.L = L
        cfg = _lib.SdnsConfig()
        cfg.abi_version = 1
        for i in range(3):
            cfg.N[i], cfg.L[i] = int(N[i]), float(Lbox[i])
            cfg.kcut[i] = -1 if kcut is None else int(kcut[i])
        cfg.precision = 1 if precision == 'double' else 0
        cfg.dealias = _lib.DEALIAS[dealias]
        cfg.solver = _lib.SOLVER[solver]
        conv = convection or ('Divergence' if solver == 'MHD' else 'Vortex')
        cfg.convection = _lib.CONVECTION[conv]
        cfg.mask_nyquist = int(mask_nyquist)
        cfg.rank, cfg.nranks = rank, nranks
        cfg.k1_layout = _lib.K1_LAYOUT[k1_layout]
        self.p = vp()
        self.chk(L.sdns_plan_create(C.byref(self.p), C.byref(cfg)))
        self.real = np.float64 if precision == 'double' else np.float32
        self.cplx = np.complex128 if precision == 'double' else np.complex64
        sp, ph, pd = (C.c_int32*3)(), (C.c_int32*3)(), (C.c_int32*3)()
        self.chk(L.sdns_local_shapes(self.p, C.byref(sp), C.byref(ph), C.byref(pd)))
        self.sshape, self.pshape, self.dshape = tuple(sp), tuple(ph), tuple(pd)
        a, b = C.c_int32(), C.c_int32()
        self.chk(L.sdns_k1_layout(self.p, C.byref(a), C.byref(b)))
        self.k1_slice = slice(a.value, int(N[1]), b.value) if b.value > 1 else slice(a.value, a.value + sp[1])     # this rank's axis-1 modes
        self.ncomp = 6 if solver == 'MHD' else 3
        self.rank, self.nranks = rank, nranks

        def share(m):                       # mpi4py-fft's slabs: m // P per rank, the first m % P ranks one more
            c, rem = divmod(m, nranks)
            lo = rank*c + min(rank, rem)
            return slice(lo, lo + c + (1 if rank < rem else 0))
        self.x0_slice, self.x0p_slice = share(int(N[0])), share(int(N[0])*3//2 if dealias == '3/2-rule' else int(N[0]))
        assert self.pshape[0] == len(range(*self.x0_slice.indices(int(N[0])))), (self.pshape, self.x0_slice)
        if nranks > 1:
            self.chk(L.sdns_comm_alloc(self.p))
        if nranks == 1:
            n = C.c_size_t()
            self.chk(L.sdns_workspace_bytes(self.p, C.byref(n)))
            self._ws = np.full(n.value + 512, 0xFF, dtype=np.uint8)      # NaN bytes: uninitialised reads show up
            base = self._ws.ctypes.data
            self.chk(L.sdns_plan_set_workspace(self.p, vp(base + (-base) % 256), n.value))

    def handle(self):
        h = (C.c_char*64)()
        self.chk(self.L.sdns_comm_handle(self.p, h))
        return bytes(h)

    def open_peers(self, handles):
        buf = (C.c_char*(64*self.nranks))(*b''.join(handles))
        self.chk(self.L.sdns_comm_open(self.p, buf, self.nranks))

    def timed_out(self):
        v = C.c_int()
        self.chk(self.L.sdns_comm_status(self.p, C.byref(v)))
        return bool(v.value)

    def chk(self, rc):
        if rc:
            raise RuntimeError('libsdns_emu: %s (%d)' % (self.L.sdns_last_error().decode(), rc))

    def close(self):
        if self.p:
            self.L.sdns_plan_destroy(self.p)
            self.p = None

    def spectral(self, nc=None):
        return np.zeros((nc or self.ncomp,) + self.sshape, dtype=self.cplx)

    def forward(self, u, padded=False):
        u = np.ascontiguousarray(u, dtype=self.real)
        out = self.spectral(u.shape[0])
        self.chk(self.L.sdns_forward(self.p, 1 if padded else 0, u.shape[0], u.ctypes.data, out.ctypes.data))
        return out

    def backward(self, u_hat, padded=False):
        u_hat = np.ascontiguousarray(u_hat, dtype=self.cplx)
        out = np.zeros((u_hat.shape[0],) + (self.dshape if padded else self.pshape), dtype=self.real)
        self.chk(self.L.sdns_backward(self.p, 1 if padded else 0, u_hat.shape[0], u_hat.ctypes.data, out.ctypes.data))
        return out

    def compute_rhs(self, u_hat, nu, eta=0.0):
        u_hat = np.ascontiguousarray(u_hat, dtype=self.cplx)
        rhs = np.zeros_like(u_hat)
        self.chk(self.L.sdns_compute_rhs(self.p, rhs.ctypes.data, u_hat.ctypes.data, nu, eta, None, None))
        return rhs

    def rk4(self, u_hat, nsteps, dt, nu, eta=0.0):
        u = np.ascontiguousarray(u_hat, dtype=self.cplx).copy()
        u1, u2 = np.zeros_like(u), np.zeros_like(u)
        for _ in range(nsteps):
            self.chk(self.L.sdns_rk4_step(self.p, u.ctypes.data, u1.ctypes.data, u2.ctypes.data, dt, nu, eta, None))
        return u


def run_ranks(world, fn):
    """Run fn(rank, sync) on `world` threads (the ranks of an emulated multi-GPU run; ctypes releases the GIL inside the
    library, so the flag barriers between the ranks make progress).  sync(obj) all-gathers obj across the ranks."""
    import threading
    bar = threading.Barrier(world)
    box = [None]*world
    out, errs = [None]*world, []

    def worker(r):
        def sync(obj):
            box[r] = obj
            bar.wait()
            got = list(box)
            bar.wait()
            return got
        try:
            out[r] = fn(r, sync)
        except BaseException as e:      # noqa: BLE001
            errs.append((r, e))
            bar.abort()
    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0][1]
    return out
