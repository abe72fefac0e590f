"""Host-side logic of the lazy state mirror (spectraldns_b200/spaces.py, SDNS_LAZY_STATE=1) without a GPU: the
DeviceState is given a stand-in plan whose "device" tensors are torch CPU tensors and whose reductions are numpy.
Checks that the per-step expressions of demo/Isotropic.py's update() (:161-184) are answered without copying the
state, that every other access to the array refreshes the mirror first, and that host writes reach the device."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import sdns_oracle as so                                    # noqa: E402
from spectraldns_b200 import spaces                         # noqa: E402
from spectraldns_b200.device_state import DeviceState       # noqa: E402


class FakePlan(object):
    """What DeviceState and the hooks use of Plan, on CPU tensors (test stand-in; the product has no CPU path)."""
    def __init__(self, o):
        self.o = o
        self.spectral_shape = tuple(o.sshape)
        self.precision, self.nranks, self.rank = 'double', 1, 0
        self.device = torch.device('cpu')
        self.tcomplex, self.complex = torch.complex128, np.dtype(np.complex128)
        self.calls = []

    def empty_spectral(self, ncomp=3):
        return torch.zeros((ncomp,)+self.spectral_shape, dtype=self.tcomplex)

    def use_current_stream(self):
        pass

    def energy_weighted(self, u, w=None):
        self.calls.append('energy_weighted')
        a = u.numpy()
        return self.o.energy_fourier(a if w is None else a*w.numpy())

    def scale_field(self, u, f, a=1.0, b=0.0):
        self.calls.append('scale_field')
        fn = f.numpy()
        u.mul_(torch.from_numpy(a*fn + b*(1-fn)))
        return u

    def set_mode(self, u, idx, value=0.0):
        self.calls.append('set_mode')
        u[(slice(None),)+tuple(idx)] = complex(value)
        return u


class FakeEngine(object):
    def __init__(self, plan):
        self.plan = plan


class FakeSpace(object):
    def __init__(self, o):
        self.o, self.complex, self.float, self.comm = o, np.dtype(np.complex128), np.dtype(np.float64), None

    def shape(self, forward_output=False):
        return tuple(self.o.sshape) if forward_output else tuple(self.o.N)


class FakeVector(spaces.CompositeSpace):
    def __init__(self, T):
        spaces.CompositeSpace.__init__(self, [T]*3)


@pytest.fixture
def lazy_setup():
    o = so.Oracle((16, 16, 16))
    T = FakeSpace(o)
    VT = FakeVector(T)
    U_hat = spaces.Function(VT)
    U_hat[...] = so.isotropic_field(o, seed=2)
    plan = FakePlan(o)
    dev = DeviceState(FakeEngine(plan), U_hat, 3)
    dev.begin_solve(lazy=True)
    yield o, T, U_hat, plan, dev
    dev.end_solve()
    assert dev not in spaces._LAZY


def test_isotropic_update_expressions_stay_on_the_device(lazy_setup):
    o, T, U_hat, plan, dev = lazy_setup
    k2_mask = np.where(o.K2 <= 3**2, 1, 0)
    target = o.energy_fourier(np.array(U_hat))
    ref = np.array(U_hat)
    dev.upload_state()
    h2d0 = dev.h2d_copies
    # one "time step" on the device: the mirror is now stale
    dev.u.mul_(0.97)
    dev.device_newer = True
    ref *= 0.97
    # demo/Isotropic.py:162-184, verbatim expressions
    U_hat[:, 0, 0, 0] = 0
    energy_new = spaces.energy_fourier(U_hat, T)
    energy_lower = spaces.energy_fourier(U_hat*k2_mask, T)
    energy_upper = energy_new - energy_lower
    alpha = np.sqrt((target - energy_upper)/energy_lower)
    U_hat *= (alpha*k2_mask + (1-k2_mask))
    energy_after = spaces.energy_fourier(U_hat, T)
    assert getattr(dev, 'd2h_copies', 0) == 0 and dev.h2d_copies == h2d0          # the state never crossed
    assert plan.calls == ['set_mode', 'energy_weighted', 'energy_weighted', 'scale_field', 'energy_weighted']
    ref, e_ref, e_low_ref, a_ref = o.forcing_rescale(ref, 3, target)
    assert abs(energy_after - e_ref) < 1e-13*abs(e_ref) and abs(alpha - a_ref) < 1e-13
    assert abs(energy_after - target) < 1e-7
    # a host read refreshes the mirror (one D2H) and sees the forced field
    got = U_hat[1, 2, 3, 4]
    assert dev.d2h_copies == 1
    assert got == ref[1, 2, 3, 4]
    assert np.allclose(np.asarray(U_hat.view(np.ndarray)), ref, rtol=1e-14, atol=0)


def test_host_access_paths_refresh_and_mark_dirty(lazy_setup):
    o, T, U_hat, plan, dev = lazy_setup
    dev.upload_state()
    base = np.array(U_hat.view(np.ndarray))
    dev.u.mul_(2.0)
    dev.device_newer = True
    s = np.sum(U_hat)                       # numpy function on the stale mirror: refreshed first
    assert np.isclose(s, 2*base.sum()) and dev.d2h_copies == 1
    dev.u.mul_(0.5)
    dev.device_newer = True
    v = abs(U_hat)                          # ufunc
    assert np.allclose(v, np.abs(base)) and dev.d2h_copies == 2
    dev.u.mul_(3.0)
    dev.device_newer = True
    c = U_hat.copy()                        # method that reads the buffer directly
    assert np.allclose(c, 3*base) and dev.d2h_copies == 3
    # a host write through a slice: mirror refreshed, then marked dirty so that the next step uploads it
    dev.u.mul_(2.0)
    dev.device_newer = True
    dev.host_dirty = False
    U_hat[0] = 1.0
    assert dev.d2h_copies == 4 and dev.host_dirty
    assert np.allclose(np.asarray(U_hat.view(np.ndarray))[1], 6*base[1])
    dev.upload_state()
    assert torch.all(dev.u[0] == 1.0)
    # a product that is used for something other than energy_fourier is just the product
    k2_mask = np.where(o.K2 <= 3**2, 1, 0)
    prod = U_hat*k2_mask
    assert isinstance(prod, spaces._ScaledState)
    assert np.allclose(np.asarray(prod), np.asarray(U_hat.view(np.ndarray))*k2_mask)
    assert np.allclose(prod[2], np.asarray(U_hat.view(np.ndarray))[2]*k2_mask)


def test_eager_mode_is_untouched():
    o = so.Oracle((8, 8, 8))
    T = FakeSpace(o)
    U_hat = spaces.Function(FakeVector(T))
    U_hat[...] = 1 + 2j
    assert not spaces._LAZY
    k = np.ones(o.sshape)
    p = U_hat*k
    assert isinstance(p, spaces.Function) and p._space is U_hat._space and np.all(p == 1 + 2j)
    U_hat *= 2
    assert np.all(U_hat == 2 + 4j) and isinstance(2*U_hat, spaces.Function)
    assert np.sum(U_hat) == U_hat.size*(2 + 4j)
