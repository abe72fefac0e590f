"""Kernel logic without a GPU: the library's own CUDA sources (plan, C ABI, every pass kernel) compiled by g++
against tests/host/host_shim.h -- one OS thread per CUDA thread, barriers for __syncthreads / __syncwarp, mailboxes
for warp shuffles, NaN-filled shared memory and workspace -- and run on tiny grids through the same C ABI calls as
the GPU parity tests, against the CPU oracle.  This checks indexing, axis maps, truncation / padding, exchange maps,
epilogues and stage updates of the PRODUCT kernels; speed and anything that depends on real hardware (memory model,
occupancy) stay with the -m gpu tests.  The emulated library is test infrastructure: the product never loads it."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'tests', 'host'), os.path.join(ROOT, 'oracle')]
import sdns_oracle as so          # noqa: E402
from conftest import rel_l2      # noqa: E402

TOL = {'double': 1e-11, 'single': 1e-4}


@pytest.fixture(scope='module')
def emu():
    import build_emu
    import emu_plan
    return emu_plan.load(build_emu.build()), emu_plan


def _state(o, solver, seed=3):
    f0 = so.isotropic_field(o, seed=seed, ncomp=6 if solver == 'MHD' else 3)
    if solver == 'VV':
        f0 = o.cross2(o.K, f0)
    return f0.astype(o.complex)


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('N', [(16, 16, 16), (8, 12, 24), (32, 16, 8), (24, 48, 16)])
def test_emulated_plain_transforms(emu, N, precision):
    L, ep = emu
    o = so.Oracle(N, precision=precision, dealias='None')
    p = ep.EmuPlan(L, N, precision=precision, dealias='None')
    rng = np.random.RandomState(1)
    u = rng.standard_normal((3,)+tuple(N)).astype(o.float)
    assert rel_l2(p.forward(u), o.forward(u)) < TOL[precision]
    assert rel_l2(p.backward(o.forward(u).astype(o.complex)), u) < TOL[precision]
    p.close()


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('dealias', ['2/3-rule', '3/2-rule', 'None'])
@pytest.mark.parametrize('solver,N', [('NS', (16, 16, 16)), ('NS', (32, 16, 8)), ('VV', (16, 32, 16)), ('MHD', (16, 16, 32))])
def test_emulated_rhs_and_rk4(emu, solver, N, dealias, precision):
    L, ep = emu
    o = so.Oracle(N, precision=precision, dealias=dealias)
    p = ep.EmuPlan(L, N, precision=precision, dealias=dealias, solver=solver)
    f0 = _state(o, solver)
    nu, eta, dt = 0.005, 0.01, 0.002
    ref = {'NS': lambda: o.ns_rhs(f0, nu), 'VV': lambda: o.vv_rhs(f0, nu), 'MHD': lambda: o.mhd_rhs(f0, nu, eta)}[solver]()
    assert rel_l2(p.compute_rhs(f0, nu, eta), ref) < TOL[precision]
    got = p.rk4(f0, 2, dt, nu, eta)
    assert rel_l2(got, o.solve(f0, solver, 2, dt, nu, eta=eta)) < TOL[precision]
    p.close()


@pytest.mark.parametrize('conv', ['Standard', 'Divergence', 'Skewed'])
def test_emulated_ns_convection_forms(emu, conv):
    L, ep = emu
    N = (16, 16, 16)
    o = so.Oracle(N)
    p = ep.EmuPlan(L, N, convection=conv)
    f0 = _state(o, 'NS')
    assert rel_l2(p.compute_rhs(f0, 0.005), o.ns_rhs(f0, 0.005, conv)) < 1e-11
    p.close()
