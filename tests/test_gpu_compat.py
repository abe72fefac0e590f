"""GPU tests of the drop-in boundary: the reference's solver-module interface (get_solver /
get_context / solve / ComputeRHS / callbacks) on the B200 implementation, driven the way the
reference's own tests drive it (tests/test_NSVV.py, tests/test_MHD.py, tests/TG.py, tests/TGMHD.py,
demo/Isotropic.py of the reference)."""
import glob
import os
import subprocess
import sys
import numpy as np
import pytest
from conftest import rel_l2, golden
import sdns_oracle as so

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, 'spectraldns_b200', 'compat')


@pytest.fixture(scope='module')
def sdns():
    sys.path.insert(0, ROOT)
    from spectraldns_b200 import run
    run.activate()
    import spectralDNS
    return spectralDNS


def tg_initialize(solver, context, config):
    """tests/TG.py:15-36 of the reference."""
    U, X = context.U, context.X
    U[0] = np.sin(X[0])*np.cos(X[1])*np.cos(X[2])
    U[1] = -np.cos(X[0])*np.sin(X[1])*np.cos(X[2])
    U[2] = 0
    solver.set_velocity(**context)
    if 'NS' not in config.params.solver:
        solver.cross2(context.W_hat, context.K, context.U_hat)
    config.params.t = 0.0
    config.params.tstep = 0


def make_tg_regression(config, store):
    def regression_test(context):
        """tests/TG.py:116-126 of the reference."""
        params, solver = config.params, config.solver
        U = solver.get_velocity(**context)
        curl = solver.get_curl(**context)
        w = solver.comm.reduce(np.sum(curl.astype(np.float64)*curl.astype(np.float64))/np.prod(params.N)/2)
        k = solver.comm.reduce(np.sum(U.astype(np.float64)*U.astype(np.float64))/np.prod(params.N)/2)
        store['k'], store['w'] = float(k), float(w)
        assert round(float(w) - 0.375249930801, params.ntol) == 0, w
        assert round(float(k) - 0.124953117517, params.ntol) == 0, k
    return regression_test


MESH = {'uniform': ['--M', '4', '4', '4', '--L', '2*pi', '2*pi', '2*pi'],
        'nonuniform': ['--M', '6', '5', '4', '--L', '6*pi', '4*pi', '2*pi']}


@pytest.mark.parametrize('mesh', ['uniform', 'nonuniform'])
@pytest.mark.parametrize('name', ['NS', 'VV'])
def test_tg_solvers(sdns, name, mesh, tmp_path, monkeypatch):
    """tests/test_NSVV.py:36-71 of the reference (test_solvers)."""
    monkeypatch.chdir(tmp_path)
    config, get_solver, solve = sdns.config, sdns.get_solver, sdns.solve
    config.update({'nu': 0.000625, 'dt': 0.01, 'T': 0.1, 'convection': 'Vortex'})
    store = {}
    solver = get_solver(regression_test=make_tg_regression(config, store), parse_args=MESH[mesh]+[name])
    context = solver.get_context()
    tg_initialize(solver, context, config)
    solve(solver, context)
    assert 'k' in store
    # field-level parity with the fixture written by the unmodified reference
    g = golden('tg_%s_%s_double' % (name.lower(), '16' if mesh == 'uniform' else '64x32x16'))
    assert rel_l2(np.array(context.u), g['u_hat']) < 1e-11
    # results + checkpoint files (test_NSVV.py:64-71)
    config.params.write_result = 2
    config.params.checkpoint = 2
    config.params.t, config.params.tstep, config.params.T = 0.0, 0, 0.04
    solver.regression_test = lambda c: None
    solve(solver, context)
    z = np.load(name + '_c.npz')
    assert int(z['attr__tstep']) == 4
    key = 'U/3D/0' if name == 'NS' else 'curl/3D/0'
    assert rel_l2(z[key], np.array(context.u)) == 0.0
    assert os.path.exists(name + '_w_t4.npz')
    config.params.write_result = config.params.checkpoint = 1e8
    config.params.T = 0.1


@pytest.mark.parametrize('integrator,ntol', [('RK4', 7), ('ForwardEuler', 4), ('AB2', 4), ('BS5_adaptive', 7), ('BS5_fixed', 7)])
def test_integrators(sdns, integrator, ntol):
    """tests/test_NSVV.py:73-92 of the reference (the explicit fixed-step integrators)."""
    config, get_solver, solve = sdns.config, sdns.get_solver, sdns.solve
    config.update({'nu': 0.000625, 'dt': 0.01, 'T': 0.1, 'convection': 'Vortex'})
    store = {}
    solver = get_solver(regression_test=make_tg_regression(config, store), parse_args=MESH['uniform']+['NS'])
    context = solver.get_context()
    config.params.ntol = ntol
    config.params.integrator = integrator
    tg_initialize(solver, context, config)
    solve(solver, context)
    config.params.ntol, config.params.integrator = 7, 'RK4'


@pytest.mark.parametrize('integrator', ['ForwardEuler', 'AB2', 'BS5_fixed', 'BS5_adaptive'])
def test_integrator_fields_match_reference(sdns, integrator):
    """Field-level parity of every explicit integrator with the reference's own
    (maths/integrators.py:15-175) on a broadband field, fixtures integ_ns_16_*."""
    config, get_solver, solve = sdns.config, sdns.get_solver, sdns.solve
    g = golden('integ_ns_16_' + integrator.lower())
    config.update({'nu': float(g['nu']), 'dt': float(g['dt']), 'T': float(g['T']), 'convection': 'Vortex'})
    solver = get_solver(regression_test=lambda c: None, parse_args=MESH['uniform'] + ['--integrator', integrator, 'NS'])
    c = solver.get_context()
    c.U_hat[:] = g['u0_hat']
    config.params.t, config.params.tstep = 0.0, 0
    solve(solver, c)
    assert config.params.tstep == int(g['nsteps'])
    assert abs(config.params.t - float(g['t_end'])) < 1e-12
    assert rel_l2(np.array(c.U_hat), g['u_hat']) < 1e-10
    config.update({'nu': 0.000625, 'dt': 0.01, 'T': 0.1})
    config.params.integrator = 'RK4'


def test_update_callback_sees_current_fields(sdns):
    """tests/TG.py:42-109 of the reference: a user update() reading velocity, curl, pressure and
    checking Parseval / divergence every other step -- exercises the device->host refresh."""
    config, get_solver, solve = sdns.config, sdns.get_solver, sdns.solve
    from shenfun.fourier import energy_fourier
    config.update({'nu': 0.000625, 'dt': 0.01, 'T': 0.1, 'convection': 'Vortex'})
    seen = []

    def update(context):
        params, solver = config.params, config.solver
        if params.tstep % 2 == 0:
            U = solver.get_velocity(**context)
            solver.get_curl(**context)
            solver.get_pressure(**context)
            kk = np.sum(U.astype(np.float64)**2)/np.prod(params.N)/2
            ww2 = energy_fourier(context.U_hat, context.T)/2
            divu = solver.get_divergence(**context)
            seen.append((params.tstep, kk, ww2 - kk, float(np.sum(divu.astype(np.float64)**2))))

    store = {}
    solver = get_solver(update=update, regression_test=make_tg_regression(config, store),
                        parse_args=['--M', '5', '5', '5', 'NS'])
    context = solver.get_context()
    tg_initialize(solver, context, config)
    solve(solver, context)
    assert [s[0] for s in seen] == [2, 4, 6, 8, 10]
    ks = [s[1] for s in seen]
    assert all(ks[i+1] < ks[i] for i in range(4))             # energy decays
    assert abs(ks[-1] - 0.124953117517) < 1e-9
    assert max(abs(s[2]) for s in seen) < 1e-14               # Parseval
    assert max(s[3] for s in seen) < 1e-20                    # divergence free


@pytest.mark.parametrize('mesh', ['uniform', 'nonuniform'])
def test_mhd(sdns, mesh):
    """tests/test_MHD.py:32-50 with tests/TGMHD.py:4-26 of the reference."""
    config, get_solver, solve = sdns.config, sdns.get_solver, sdns.solve
    config.update({'nu': 0.000625, 'dt': 0.01, 'T': 0.1, 'eta': 0.01, 'convection': 'Divergence'})
    res = {}

    def regression_test(context):
        params, solver = config.params, config.solver
        dx, L = params.dx, params.L
        UB = context.UB_hat.backward(context.UB)
        U, B = UB[:3], UB[3:]
        k = np.sum(U.astype(np.float64)**2)*dx[0]*dx[1]*dx[2]/L[0]/L[1]/L[2]/2
        b = np.sum(B.astype(np.float64)**2)*dx[0]*dx[1]*dx[2]/L[0]/L[1]/L[2]/2
        res['k'], res['b'] = float(k), float(b)
        assert round(float(k) - 0.124565408177, 7) == 0
        assert round(float(b) - 0.124637762143, 7) == 0

    solver = get_solver(regression_test=regression_test, parse_args=MESH[mesh]+['MHD'])
    c = solver.get_context()
    U, B, X = c.U, c.B, c.X
    U[0] = np.sin(X[0])*np.cos(X[1])*np.cos(X[2])
    U[1] = -np.cos(X[0])*np.sin(X[1])*np.cos(X[2])
    U[2] = 0
    B[0] = np.sin(X[0])*np.sin(X[1])*np.cos(X[2])
    B[1] = np.cos(X[0])*np.cos(X[1])*np.cos(X[2])
    B[2] = 0
    c.UB.forward(c.UB_hat)
    config.params.t, config.params.tstep = 0, 0
    solve(solver, c)
    assert 'b' in res
    g = golden('tg_mhd_%s_double' % ('16' if mesh == 'uniform' else '64x32x16'))
    assert rel_l2(np.array(c.UB_hat), g['u_hat']) < 1e-11
    config.update({'convection': 'Vortex'})


@pytest.mark.parametrize('lazy', [False, True])
@pytest.mark.parametrize('precision,dealias', [('double', '3/2-rule'), ('single', '3/2-rule'), ('double', '2/3-rule')])
def test_forced_isotropic_callback_parity(sdns, precision, dealias, lazy, monkeypatch):
    """demo/Isotropic.py:29-76,150-187 of the reference: broadband initial field, and an update()
    that rescales the low wavenumber band of c.U_hat every step (the forcing), written with numpy on the
    context's host array.  The same sequence is replayed with the CPU oracle; relative L2 of the velocity
    spectrum.  lazy: SDNS_LAZY_STATE=1 -- the same callback, but the state never crosses PCIe inside the
    time loop (its expressions are answered by sdns_energy_weighted / sdns_scale_field / sdns_set_mode)."""
    monkeypatch.setenv('SDNS_LAZY_STATE', '1' if lazy else '0')
    config, get_solver, solve = sdns.config, sdns.get_solver, sdns.solve
    from shenfun.fourier import energy_fourier
    config.update({'nu': 0.005428, 'dt': 0.002, 'T': 0.01, 'convection': 'Vortex'})
    N = (32, 32, 32)
    o = so.Oracle(N, precision=precision, dealias=dealias)
    u0 = so.isotropic_field(o, seed=7)
    k2_mask = np.where(o.K2 <= 3**2, 1, 0)
    target = o.energy_fourier(u0)
    log = []

    def forcing(U_hat, efun):
        energy_new = efun(U_hat)
        energy_lower = efun(U_hat*k2_mask)
        alpha = np.sqrt((target - (energy_new - energy_lower))/energy_lower)
        U_hat *= (alpha*k2_mask + (1-k2_mask))
        return efun(U_hat)

    def update(c):
        c.U_hat[:, 0, 0, 0] = 0
        e = forcing(c.U_hat, lambda a: energy_fourier(a, c.T))
        log.append(e)
        assert abs(e - target) < (1e-7 if precision == 'double' else 1e-4)     # demo/Isotropic.py:184

    args = ['--M', '5', '5', '5', '--precision', precision, '--dealias', dealias, 'NS']
    solver = get_solver(update=update, regression_test=lambda c: None, parse_args=args)
    nu = float(config.params.nu)
    c = solver.get_context()
    c.U_hat[:] = u0
    config.params.t, config.params.tstep = 0.0, 0
    solve(solver, c)
    assert len(log) == 5
    dev = c['_dev']
    if lazy:
        # one upload of the initial field, one download when solve() hands the result back: nothing per step
        assert dev.h2d_copies == 1 and dev.d2h_copies == 1, (dev.h2d_copies, dev.d2h_copies)
    else:
        assert dev.d2h_copies >= 5 and dev.h2d_copies >= 5
    # oracle replay
    u = u0.copy()
    for _ in range(5):
        u = o.rk4_step(u, lambda v: o.ns_rhs(v, nu), config.params.dt)
        u[:, 0, 0, 0] = 0
        forcing(u, o.energy_fourier)
    tol = 1e-11 if precision == 'double' else 1e-4
    assert rel_l2(np.array(c.U_hat), u) < tol
    # direct ComputeRHS on the context arrays (demo/Isotropic.py:233)
    dU = solver.ComputeRHS(c.dU, c.U_hat, solver, **c)
    assert rel_l2(np.array(dU), o.ns_rhs(u, nu)) < 10*tol
    config.update({'nu': 0.000625, 'dt': 0.01, 'T': 0.1})
    config.params.dealias, config.params.precision = '2/3-rule', 'double'


def _run(cmd, cwd, extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    env['PYTHONPATH'] = COMPAT + os.pathsep + ROOT + os.pathsep + env.get('PYTHONPATH', '')
    r = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    return r.returncode, r.stdout


def test_reference_scripts_run_unchanged(tmp_path):
    """The reference's own demo/test drivers, byte-for-byte (staged by oracle/stage_reference_scripts.sh
    under baseline/_ref, which travels to the GPU box), on top of the B200 `spectralDNS`."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.exists(os.path.join(ref, 'tests', 'TG.py')):
        pytest.skip('reference scripts not staged (run oracle/stage_reference_scripts.sh in the build container)')
    py = sys.executable
    for script, args in (('tests/TG.py', ['NS']), ('tests/TG.py', ['VV']), ('demo/TG.py', ['--M', '6', '6', '6', 'NS']),
                         ('tests/TGMHD.py', ['MHD']), ('demo/TGMHD.py', ['MHD'])):
        rc, out = _run([py, os.path.join(ref, script)] + args, str(tmp_path))
        assert rc == 0, (script, args, out[-3000:])
        assert 'Fastest' in out
    iso = [py, os.path.join(ref, 'demo', 'Isotropic.py'), '--N', '32', '32', '32', '--T', '0.02', '--compute_energy', '5', 'NS']
    rc, out = _run(iso, str(tmp_path))
    assert rc == 0, out[-3000:]
    # the same unchanged script with the lazily mirrored state: same monitor lines (t, energies, dissipation, Re_lambda)
    rc, out_lazy = _run(iso, str(tmp_path), {'SDNS_LAZY_STATE': '1'})
    assert rc == 0, out_lazy[-3000:]

    def monitor(text):
        rows = []
        for line in text.splitlines():
            f = line.split()
            if len(f) == 9:
                try:
                    rows.append([float(x) for x in f])
                except ValueError:
                    pass
        return np.array(rows)
    m0, m1 = monitor(out), monitor(out_lazy)
    assert m0.shape == m1.shape and m0.shape[0] >= 1
    assert np.allclose(m0, m1, rtol=1e-5, atol=1e-9), (m0, m1)
    rc, out = _run([py, '-m', 'pytest', '-x', '-q', os.path.join(ref, 'tests', 'test_NSVV.py'),
                    os.path.join(ref, 'tests', 'test_MHD.py'), '-p', 'no:cacheprovider'], str(tmp_path))
    assert rc == 0, out[-4000:]
