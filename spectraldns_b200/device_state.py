"""Device-resident solver state behind the reference's host-array interface.

The reference's context holds numpy arrays that integrate() updates in place and that user
callbacks read and write (demo/Isotropic.py:161-187).  Here the state lives on the GPU while the
time loop runs; the host arrays in the context are mirrors.  Rules:

  * integrate() uploads the host state only if the host may have changed since the last
    upload, advances the device state, and marks the host mirror stale.
  * solve() brings the mirror up to date (one D2H copy) before it calls any user-installed callback
    or writes a file, and afterwards assumes the host may have been modified.
  * With the default no-op callbacks nothing is copied inside the time loop.
  * Outside solve() (user code calling integrate()/ComputeRHS directly) every call is eager:
    upload before, download after.
"""
import numpy as np


class DeviceState(object):
    def __init__(self, engine, host_state, ncomp):
        import torch
        self.torch = torch
        self.engine = engine
        self.plan = engine.plan
        self.host_state = host_state          # context.u (numpy, ideally pinned)
        p = self.plan
        self.u = p.empty_spectral(ncomp)
        self.u1 = p.empty_spectral(ncomp)
        self.u2 = p.empty_spectral(ncomp)
        self.rhs = None
        self.source = None
        self.ncomp = ncomp
        self.host_dirty = True                # host may be newer than the device copy
        self.device_newer = False             # host mirror is stale
        self.managed = False                  # True while solve() owns the sync points
        self.source_active = False

    # -- copies -------------------------------------------------------------
    def _h2d(self, dst, src):
        src = np.asarray(src).view(np.ndarray) if isinstance(src, np.ndarray) else src
        t = self.torch.from_numpy(np.ascontiguousarray(src))
        dst.copy_(t.reshape(dst.shape), non_blocking=False)

    def _d2h(self, dst, src):
        dst = dst.view(np.ndarray)            # plain view of the same memory: no array-subclass hooks on the way
        t = self.torch.from_numpy(dst) if dst.flags['C_CONTIGUOUS'] else None
        if t is not None:
            t.reshape(src.shape).copy_(src)
        else:
            dst[...] = src.cpu().numpy().reshape(dst.shape)

    def upload_state(self, host=None):
        h = self.host_state if host is None else host
        self._h2d(self.u, h)
        self.h2d_copies = getattr(self, 'h2d_copies', 0) + 1
        if host is None or host is self.host_state:
            self.host_dirty = False
            self.device_newer = False

    def sync_to_host(self):
        if self.device_newer:
            self.device_newer = False         # first: the copy below may pass through the lazy-mirror hooks of the array
            if self.plan.nranks > 1:
                self.plan.sync()              # raises SdnsError if a cross-GPU barrier timed out: never hand out garbage
            self._d2h(self.host_state, self.u)
            self.d2h_copies = getattr(self, 'd2h_copies', 0) + 1

    def is_state(self, a):
        h = self.host_state
        return a is h or (isinstance(a, np.ndarray) and a.shape == h.shape and a.dtype == h.dtype
                          and a.ctypes.data == h.ctypes.data)

    def device_input(self, host_array):
        """Device tensor holding `host_array`'s values: the resident state when it is the
        registered (and current) state array, else a staged upload."""
        if self.is_state(host_array):
            if self.host_dirty or not self.managed:
                self.upload_state()
            return self.u
        p = self.plan
        return self.engine.upload('rhs_in', np.asarray(host_array).reshape(self.u.shape), p.complex, p.tcomplex)

    def rhs_buffer(self):
        if self.rhs is None:
            self.rhs = self.plan.empty_spectral(self.ncomp)
        return self.rhs

    def refresh_source(self, host_source):
        """Upload Source when it is non-zero (NS.py:259, VV.py:108); None otherwise."""
        if host_source is None:
            self.source_active = False
            return None
        a = np.asarray(host_source)
        if not a.any():
            self.source_active = False
            return None
        if self.source is None:
            self.source = self.plan.empty_spectral(self.ncomp)
        self._h2d(self.source, a.astype(self.plan.complex, copy=False))
        self.source_active = True
        return self.source


def _solve_hooks():
    def begin_solve(self, lazy=False):
        self.managed = True
        self.host_dirty = True
        self.lazy = bool(lazy)
        if self.lazy:
            from .spaces import _LAZY
            if self not in _LAZY:
                _LAZY.append(self)

    def end_solve(self):
        self.sync_to_host()
        if getattr(self, 'lazy', False):
            from .spaces import _LAZY
            if self in _LAZY:
                _LAZY.remove(self)
            self.lazy = False
        self.managed = False
        self.host_dirty = True

    def host_touched(self):
        self.host_dirty = True

    DeviceState.begin_solve = begin_solve
    DeviceState.end_solve = end_solve
    DeviceState.host_touched = host_touched


_solve_hooks()
