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
