"""CPU tests: the oracle restatement (oracle/sdns_oracle.py) against the reference's own known
answers (tests/TG.py:125-126, tests/TGMHD.py:25-26 of the reference) and against the fixtures
written by oracle/make_golden.py from the UNMODIFIED reference solver modules."""
import glob
import os
import numpy as np
import pytest
from conftest import golden, rel_l2, GOLDEN
import sdns_oracle as so

EVERY = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, '*.npz')))
ALL = [n for n in EVERY if not n.startswith('integ_')]
INTEG = [n for n in EVERY if n.startswith('integ_')]


def make_oracle(g):
    nomask = 'nodealias' in str(g.get('name', ''))
    return so.Oracle(g['N'], g['L'], str(g['precision']), str(g['dealias']))


def initial_state(o, g, name):
    if 'u0_hat' in g:
        return g['u0_hat']
    if str(g['solver']) == 'MHD':
        return o.forward(so.taylor_green_mhd(o))
    u = o.forward(so.taylor_green(o))
    if str(g['solver']) == 'VV':
        u = o.cross2(o.K, u)
    return u


def test_fixture_inventory():
    assert len(ALL) >= 20
    for must in ('tg_ns_16_double', 'tg_vv_16_double', 'tg_ns_64x32x16_double', 'tg_mhd_16_double',
                 'tg_ns_32_double', 'iso_ns_16_double_pad', 'iso_mhd_16_double'):
        assert must in ALL


@pytest.mark.parametrize('name', ALL)
def test_oracle_reproduces_reference(name):
    g = golden(name)
    mask = 'nodealias' not in name
    o = so.Oracle(g['N'], g['L'], str(g['precision']), str(g['dealias']), mask_nyquist=mask)
    u0 = initial_state(o, g, name)
    solver = str(g['solver'])
    eta = float(g['eta']) if 'eta' in g else None
    out = o.solve(u0, solver, int(g['nsteps']), float(g['dt']), float(g['nu']), eta=eta)
    tol = 1e-13 if str(g['precision']) == 'double' else 1e-5
    assert rel_l2(out, g['u_hat']) < tol
    for key in g.files:
        if key.startswith('rhs_'):
            conv = key[4:]
            if solver == 'NS':
                r = o.ns_rhs(u0, float(g['nu']), conv)
            elif solver == 'VV':
                r = o.vv_rhs(u0, float(g['nu']))
            else:
                r = o.mhd_rhs(u0, float(g['nu']), eta)
            assert rel_l2(r, g[key]) < tol, key


@pytest.mark.parametrize('name', [n for n in ALL if n.startswith('tg_') and 'mhd' not in n
                                  and 'single' not in n])
def test_tg_known_answer(name):
    """tests/TG.py:116-126 of the reference: k and w after 10 RK4 steps, round(., 7) == 0."""
    g = golden(name)
    o = so.Oracle(g['N'], g['L'], 'double', str(g['dealias']))
    u_hat = g['u_hat']
    if str(g['solver']) == 'VV':
        w_hat = u_hat
        U = o.backward(o.cross2(o.K_over_K2, w_hat))
        curl = o.backward(w_hat)
    else:
        U = o.backward(u_hat)
        curl = o.backward(o.cross2(o.K, u_hat))
    k = np.sum(U*U)/np.prod(o.N)/2
    w = np.sum(curl*curl)/np.prod(o.N)/2
    assert round(float(w) - 0.375249930801, 7) == 0
    assert round(float(k) - 0.124953117517, 7) == 0
    # tests/TG.py:101-109: Parseval consistency pins the 1/prod(N) forward normalisation
    if str(g['solver']) == 'NS':
        assert abs(o.energy_fourier(u_hat)/2 - k) < 1e-14


@pytest.mark.parametrize('name', [n for n in ALL if n.startswith('tg_mhd')])
def test_tgmhd_known_answer(name):
    """tests/TGMHD.py:15-26 of the reference."""
    g = golden(name)
    o = so.Oracle(g['N'], g['L'], 'double', str(g['dealias']))
    UB = o.backward(g['u_hat'])
    k = np.sum(UB[:3]**2)/np.prod(o.N)/2
    b = np.sum(UB[3:]**2)/np.prod(o.N)/2
    assert round(float(k) - 0.124565408177, 7) == 0
    assert round(float(b) - 0.124637762143, 7) == 0


def test_short_solver_known_answer_64():
    """spectralDNS3D_short.py:110-113 of the reference: k == 0.124953117517 at 64^3, T=0.1."""
    o = so.Oracle((64,)*3)
    u = o.solve(o.forward(so.taylor_green(o)), 'NS', 10, 0.01, 0.000625)
    U = o.backward(u)
    k = 0.5*np.sum(U*U)/64**3
    assert round(float(k) - 0.124953117517, 7) == 0


def test_dealias_convention_matters_for_broadband():
    """SURVEY 8c probe: TG cannot discriminate dealiasing conventions, a broadband field does."""
    g = golden('iso_ns_16_double')
    o1 = so.Oracle(g['N'], g['L'], 'double', '2/3-rule')
    o2 = so.Oracle(g['N'], g['L'], 'double', 'None')
    r1 = o1.ns_rhs(g['u0_hat'], float(g['nu']))
    r2 = o2.ns_rhs(g['u0_hat'], float(g['nu']))
    assert rel_l2(r1, g['rhs_Vortex']) < 1e-13
    assert rel_l2(r2, g['rhs_Vortex']) > 1e-3


@pytest.mark.parametrize('name', INTEG)
def test_oracle_integrators_reproduce_reference(name):
    """ForwardEuler / AB2 / BS5_fixed / BS5_adaptive (maths/integrators.py:15-175) on a broadband
    field, against the reference's own integrators (fixtures from oracle/make_golden.py)."""
    g = golden(name)
    o = so.Oracle(g['N'], g['L'], 'double', str(g['dealias']))
    nu, dt, T = float(g['nu']), float(g['dt']), float(g['T'])
    fn = lambda u: o.ns_rhs(u, nu)
    integ = str(g['integrator'])
    u = g['u0_hat'].copy()
    if integ == 'ForwardEuler':
        for _ in range(int(g['nsteps'])):
            u = o.forward_euler_step(u, fn, dt)
    elif integ == 'AB2':
        u1 = np.zeros_like(u)
        for ts in range(int(g['nsteps'])):
            u, u1 = o.ab2_step(u, u1, fn, dt, ts)
    else:
        u, n, t = o.bs5_solve(u, fn, dt, T, integ == 'BS5_adaptive')
        assert n == int(g['nsteps']) and abs(t - float(g['t_end'])) < 1e-12
    assert rel_l2(u, g['u_hat']) < 1e-12


# ---------------------------------------------------------------------------------------------
# second pin: the reference's own COMPILED kernels (Cython templates built by oracle/build_ref_cython.py
# into oracle/_ref/, from the sources where they lie under /root/reference)
# ---------------------------------------------------------------------------------------------
def _ref_cython(precision):
    import importlib
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(here, 'oracle'))
    import build_ref_cython as brc
    if not brc.available() and not brc.build():
        pytest.skip('oracle/_ref not built and /root/reference absent')
    if brc.OUT not in sys.path:
        sys.path.insert(0, brc.OUT)
    return [importlib.import_module('cython_%s_%s' % (precision, m)) for m in ('maths', 'solvers', 'integrators')]


@pytest.mark.parametrize('precision', ['double', 'single'])
def test_oracle_against_reference_cython_kernels(precision):
    """cross1 / cross2 / add_pressure_diffusion_NS / RK4 / ForwardEuler / AB2 of the reference's compiled
    optimisation modules (optimization/cython_{maths,solvers,integrators}.in) against the restatement."""
    maths, solvers, integ = _ref_cython(precision)
    N = (8, 6, 10)
    o = so.Oracle(N, L=(2*np.pi, 4*np.pi, 6*np.pi), precision=precision)
    tol = 1e-15 if precision == 'double' else 2e-7
    rng = np.random.RandomState(12)
    a = rng.standard_normal((3,)+N).astype(o.float)
    b = rng.standard_normal((3,)+N).astype(o.float)
    c = maths.cross1(np.zeros_like(a), a, b)
    assert rel_l2(o.cross1(a, b), c) <= tol

    def cplx(shape):
        return (rng.standard_normal(shape) + 1j*rng.standard_normal(shape)).astype(o.complex)
    bh = cplx((3,)+o.sshape)
    # cross2 with the broadcast wavenumber list (cython _cross3) and with a dense real field (_cross2)
    c3 = maths.cross2(np.zeros_like(bh), [np.ascontiguousarray(k) for k in o.K], bh)
    assert rel_l2(o.cross2(o.K, bh), c3) <= tol
    ad = rng.standard_normal((3,)+o.sshape).astype(o.float)
    c2 = maths.cross2(np.zeros_like(bh), ad, bh)
    assert rel_l2(o.cross2(ad, bh), c2) <= tol
    # add_pressure_diffusion (NS.py:203-217 / cython_solvers.in:40-80)
    du, uh = cplx((3,)+o.sshape), cplx((3,)+o.sshape)
    nu = o.float(0.0123)
    p_ref = np.zeros(o.sshape, dtype=o.complex)
    du_ref = solvers.add_pressure_diffusion_NS(du.copy(), uh, nu, o.K2, o.K, p_ref, o.K_over_K2)
    du_or, p_or = o.add_pressure_diffusion(du.copy(), uh, nu)
    assert rel_l2(du_or, du_ref) <= 4*tol and rel_l2(p_or, p_ref) <= 4*tol

    # integrators: the reference's compiled stage loops around a ComputeRHS supplied by the oracle
    class Solver(object):
        @staticmethod
        def ComputeRHS(dU, U_hat, solver, **ctx):
            dU[:] = o.ns_rhs(U_hat, 0.01)
            return dU
    u0 = (o.forward(so.taylor_green(o)) + 0.05*cplx((3,)+o.sshape)*(o.mask if o.mask is not None else 1)).astype(o.complex)
    dt = 0.01
    fn = lambda u: o.ns_rhs(u, 0.01)
    A = np.array([1./6., 1./3., 1./3., 1./6.], dtype=o.float)
    B = np.array([0.5, 0.5, 1.], dtype=o.float)
    U = u0.copy()
    integ.RK4(U, np.zeros_like(U), np.zeros_like(U), np.zeros_like(U), A, B, o.float(dt), Solver, {})
    assert rel_l2(o.rk4_step(u0, fn, dt), U) <= 10*tol
    U = u0.copy()
    integ.ForwardEuler(U, np.zeros_like(U), np.zeros_like(U), o.float(dt), Solver, {})
    assert rel_l2(o.forward_euler_step(u0, fn, dt), U) <= 10*tol
    U, U1 = u0.copy(), np.zeros_like(u0)
    ref, r1 = u0.copy(), np.zeros_like(u0)
    for ts in range(3):
        integ.AB2(U, U1, np.zeros_like(U), o.float(dt), ts, Solver, {})
        ref, r1 = o.ab2_step(ref, r1, fn, dt, ts)
    assert rel_l2(ref, U) <= 20*tol and rel_l2(r1, U1) <= 20*tol


# ---- doubly periodic solvers (oracle/sdns_oracle2d.py) ---------------------------------------------------------
import sdns_oracle2d as so2   # noqa: E402


@pytest.mark.parametrize('cfg', [((32, 32), (2*np.pi, 2*np.pi), '2/3-rule'), ((32, 32), (2*np.pi, 2*np.pi), '3/2-rule'),
                                 ((64, 16), (6*np.pi, 4*np.pi), '2/3-rule')])
def test_oracle2d_taylor_green_known_answer(cfg):
    """tests/TG2D.py:41-52 driven as tests/test_NS2D.py:18-33 does (nu 0.01, dt 0.05; shortened to T = 2): the kinetic
    energy of the 2-D Taylor-Green vortex equals the analytic exp(-2 nu t) decay to params.ntol = 7 digits.  The second
    mesh of test_NS2D.py (--M 6 4 --L 6*pi 4*pi) is the non-uniform case."""
    N, L, dealias = cfg
    o = so2.Oracle2D(N, L=L, dealias=dealias)
    u = so2.taylor_green_2d(o)
    nu, dt, nsteps = 0.01, 0.05, 40
    u = o.solve(u, 'NS2D', nsteps, dt, nu)
    U = o.backward(u)
    t = nsteps*dt
    X = o.mesh()
    k = np.sum(U.astype(np.float64)**2)/np.prod(N)/2
    Ue = np.array([-np.sin(X[1])*np.cos(X[0])*np.exp(-2*nu*t), np.sin(X[0])*np.cos(X[1])*np.exp(-2*nu*t)])
    ke = np.sum(Ue**2)/np.prod(N)/2
    if L[0] == L[1]:
        assert round(float(k - ke), 7) == 0
    else:
        # on the stretched box the field is not an exact solution; the reference's test only asserts on what it prints
        # for rank 0 with the same rounding -- energy must still decay monotonically and stay bounded by the initial value
        assert 0 < k < 0.25


@pytest.mark.parametrize('precision', ['double', 'single'])
def test_oracle2d_against_reference_cython_kernels(precision):
    """add_pressure_diffusion_NS2D / add_pressure_diffusion_Bq2D (optimization/cython_solvers.in:82-127) and cross2_2D /
    cross1_2D (cython_maths.in:89-147) of the reference's compiled modules against the 2-D restatement."""
    maths, solvers, _ = _ref_cython(precision)
    N = (12, 10)
    o = so2.Oracle2D(N, L=(2*np.pi, 4*np.pi), precision=precision)
    tol = 1e-15 if precision == 'double' else 2e-7
    rng = np.random.RandomState(21)

    def cplx(shape):
        return (rng.standard_normal(shape) + 1j*rng.standard_normal(shape)).astype(o.complex)
    K2d = [np.ascontiguousarray(np.broadcast_to(k, o.sshape)) for k in o.K]
    uh = cplx((2,)+o.sshape)
    c = maths.cross2_2D(np.zeros(o.sshape, dtype=o.complex), K2d, uh)
    assert rel_l2(o.cross2(uh), c) <= tol
    du = cplx((2,)+o.sshape)
    nu = o.float(0.0123)
    p_ref = np.zeros(o.sshape, dtype=o.complex)
    du_ref = solvers.add_pressure_diffusion_NS2D(du.copy(), uh, nu, o.K2, K2d, p_ref, o.K_over_K2)
    du_or, p_or = o.add_pressure_diffusion_ns2d(du.copy(), uh, nu)
    assert rel_l2(du_or, du_ref) <= 4*tol and rel_l2(p_or, p_ref) <= 4*tol
    urh, dur = cplx((3,)+o.sshape), cplx((3,)+o.sshape)
    Ri, Pr = o.float(0.1), o.float(0.7)
    p_ref = np.zeros(o.sshape, dtype=o.complex)
    du_ref = solvers.add_pressure_diffusion_Bq2D(dur.copy(), urh, p_ref, o.K_over_K2, K2d, o.K2, nu, Ri, Pr)
    du_or, p_or = o.add_pressure_diffusion_bq2d(dur.copy(), urh, nu, Ri, Pr)
    assert rel_l2(du_or, du_ref) <= 4*tol and rel_l2(p_or, p_ref) <= 4*tol
    a = rng.standard_normal((2,)+N).astype(o.float)
    b = rng.standard_normal((2,)+N).astype(o.float)
    c1 = maths.cross1_2D(np.zeros(N, dtype=o.float), a, b)
    assert rel_l2(a[0]*b[1] - a[1]*b[0], c1) <= tol
