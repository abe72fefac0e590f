"""GPU tests of the doubly periodic solvers NS2D / Bq2D (csrc/sdns2d_api.cu): parity of the
C-ABI path with the CPU oracle (oracle/sdns_oracle2d.py, pinned by tests/test_oracle.py), the reference's known answer
for the 2-D Taylor-Green vortex, and the reference's own 2-D test drivers (tests/TG2D.py, tests/test_NS2D.py) unchanged
on top of the B200 `spectralDNS`."""
import os
import subprocess
import sys
import numpy as np
import pytest
from conftest import rel_l2
import sdns_oracle2d as so2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, 'spectraldns_b200', 'compat')
TOL = {'double': 1e-11, 'single': 1e-4}      # BASELINE.json north_star tolerances (relative L2)


def make_plan(N, L=(2*np.pi,)*2, precision='double', dealias='2/3-rule', solver='NS2D'):
    from spectraldns_b200.plan import Plan2D
    return Plan2D(N, L, precision, dealias, solver)


def _state(o, ncomp, seed=5):
    rng = np.random.RandomState(seed)
    U = rng.standard_normal((ncomp,)+o.N)
    X = o.mesh()
    U[0] += 2*np.sin(X[0])*np.cos(X[1])
    U[1] -= 2*np.sin(X[1])*np.cos(X[0])
    uh = o.forward(U.astype(o.float))
    if o.mask is not None:
        uh = uh*o.mask
    return uh.astype(o.complex)


@pytest.mark.parametrize('solver,N,Lbox,dealias,precision', [
    ('NS2D', (64, 64), (2*np.pi, 2*np.pi), '2/3-rule', 'double'), ('NS2D', (64, 16), (6*np.pi, 4*np.pi), '2/3-rule', 'double'),
    ('NS2D', (32, 32), (2*np.pi, 2*np.pi), '3/2-rule', 'double'), ('NS2D', (256, 128), (2*np.pi, 2*np.pi), '3/2-rule', 'single'),
    ('NS2D', (128, 512), (2*np.pi, 4*np.pi), 'None', 'double'), ('NS2D', (1024, 1024), (2*np.pi, 2*np.pi), '2/3-rule', 'double'),
    ('NS2D', (2048, 512), (2*np.pi, 2*np.pi), '2/3-rule', 'single'),
    ('Bq2D', (64, 64), (2*np.pi, 2*np.pi), '2/3-rule', 'double'), ('Bq2D', (128, 256), (6*np.pi, 4*np.pi), '3/2-rule', 'double'),
    ('Bq2D', (256, 256), (2*np.pi, 2*np.pi), '2/3-rule', 'single'), ('Bq2D', (512, 2048), (2*np.pi, 2*np.pi), 'None', 'double')])
def test_2d_rhs_and_rk4(solver, N, Lbox, dealias, precision):
    """Transforms on T and Tp, ComputeRHS (with the pressure), RK4 / ForwardEuler / AB2 steps and the stand-alone
    operators, against the oracle on a broadband field."""
    o = so2.Oracle2D(N, L=Lbox, precision=precision, dealias=dealias)
    p = make_plan(N, Lbox, precision, dealias, solver)
    p.Ri, p.Pr = 0.1, 0.7
    tol = TOL[precision]
    nc = p.ncomp
    rng = np.random.RandomState(2)
    u = rng.standard_normal((nc,)+tuple(N)).astype(o.float)
    assert rel_l2(p.to_host(p.forward(p.to_device(u))), o.forward(u)) < tol
    assert rel_l2(p.to_host(p.backward(p.to_device(o.forward(u).astype(o.complex)))), u) < tol
    f0 = _state(o, nc)
    d_f = p.to_device(f0)
    assert rel_l2(p.to_host(p.backward(d_f, padded=True)), o._bwd_p(f0)) < tol
    up = rng.standard_normal((nc,)+tuple(o.M)).astype(o.float)
    assert rel_l2(p.to_host(p.forward(p.to_device(up), padded=True)), o._fwd_p(up)) < tol
    nu, dt = 0.01, 0.002
    if solver == 'NS2D':
        ref, pref = o.ns2d_rhs(f0, nu, return_p=True)
        fn = lambda v: o.ns2d_rhs(v, nu)
    else:
        ref, pref = o.bq2d_rhs(f0, nu, p.Ri, p.Pr, return_p=True)
        fn = lambda v: o.bq2d_rhs(v, nu, p.Ri, p.Pr)
    d_r, d_p = p.empty_spectral(), p.empty_spectral(0)
    p.compute_rhs(d_r, d_f, nu, p_hat=d_p)
    assert rel_l2(p.to_host(d_r), ref) < tol and rel_l2(p.to_host(d_p), pref) < tol
    d_u, d_1, d_2 = p.to_device(f0), p.empty_spectral(), p.empty_spectral()
    for _ in range(2):
        p.rk4_step(d_u, d_1, d_2, dt, nu)
    assert rel_l2(p.to_host(d_u), o.solve(f0, solver, 2, dt, nu, p.Ri, p.Pr)) < tol
    d_u = p.to_device(f0)
    p.euler_step(d_u, d_r, dt, nu)
    assert rel_l2(p.to_host(d_u), o.forward_euler_step(f0, fn, dt)) < tol
    d_u, d_1 = p.to_device(f0), p.empty_spectral()
    refu, r1 = f0.copy(), np.zeros_like(f0)
    for ts in range(3):
        p.ab2_step(d_u, d_1, d_r, dt, ts, nu)
        refu, r1 = o.ab2_step(refu, r1, fn, dt, ts)
    assert rel_l2(p.to_host(d_u), refu) < tol
    d_c = p.empty_spectral(0)
    assert rel_l2(p.to_host(p.cross2(d_c, d_f)), o.cross2(f0[:2])) < (1e-14 if precision == 'double' else 1e-6)
    assert p.launch_count() > 0


def test_2d_taylor_green_known_answer():
    """tests/TG2D.py:41-52 as tests/test_NS2D.py drives it (nu 0.01, dt 0.05, T 10, 2/3-rule then 3/2-rule): the
    kinetic energy equals the analytic exp(-2 nu t) decay to the reference's ntol = 7 digits."""
    for dealias in ('2/3-rule', '3/2-rule'):
        N = (32, 32)
        o = so2.Oracle2D(N, dealias=dealias)
        p = make_plan(N, dealias=dealias)
        d_u, d_1, d_2 = p.to_device(so2.taylor_green_2d(o)), p.empty_spectral(), p.empty_spectral()
        nu, dt, nsteps = 0.01, 0.05, 200
        for _ in range(nsteps):
            p.rk4_step(d_u, d_1, d_2, dt, nu)
        U = p.to_host(p.backward(d_u))
        k = np.sum(U.astype(np.float64)**2)/np.prod(N)/2
        ke = 0.25*np.exp(-2*nu*nsteps*dt)**2
        assert round(float(k - ke), 7) == 0


def _run(cmd, cwd):
    env = dict(os.environ)
    env['PYTHONPATH'] = COMPAT + os.pathsep + ROOT + os.pathsep + env.get('PYTHONPATH', '')
    r = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    return r.returncode, r.stdout


def test_reference_2d_drivers_run_unchanged(tmp_path):
    """The reference's tests/TG2D.py (as a script) and tests/test_NS2D.py (pytest: two meshes, 2/3-rule, 3/2-rule, the
    reload with --optimization cython, results + checkpoint files), byte-for-byte from baseline/_ref."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.exists(os.path.join(ref, 'tests', 'TG2D.py')):
        pytest.skip('reference scripts not staged (run oracle/stage_reference_scripts.sh in the build container)')
    py = sys.executable
    rc, out = _run([py, os.path.join(ref, 'tests', 'TG2D.py'), 'NS2D'], str(tmp_path))
    assert rc == 0, out[-3000:]
    assert 'Error' in out and 'Fastest' in out
    rc, out = _run([py, '-m', 'pytest', '-x', '-q', os.path.join(ref, 'tests', 'test_NS2D.py'), '-p', 'no:cacheprovider'],
                   str(tmp_path))
    assert rc == 0, out[-4000:]


def test_bq2d_solver_module(tmp_path, monkeypatch):
    """The Bq2D solver module through get_solver / get_context / solve (the reference has no test driver for it): a
    stratified shear layer for ten RK4 steps against the oracle, field-level."""
    monkeypatch.chdir(tmp_path)
    sys.path.insert(0, ROOT)
    from spectraldns_b200 import run
    run.activate()
    import spectralDNS
    config, get_solver, solve = spectralDNS.config, spectralDNS.get_solver, spectralDNS.solve
    config.update({'nu': 0.01, 'dt': 0.005, 'T': 0.05}, 'doublyperiodic')
    solver = get_solver(mesh='doublyperiodic', parse_args=['--M', '5', '6', '--L', '2*pi', '4*pi', 'Bq2D', '--Ri', '0.2', '--Pr', '0.7'])
    c = solver.get_context()
    N = tuple(int(n) for n in config.params.N)
    o = so2.Oracle2D(N, L=tuple(float(l) for l in config.params.L))
    X = c.X
    c.Ur[0] = np.tanh((X[1]-2*np.pi)/0.5) + 0.01*np.sin(X[0])
    c.Ur[1] = 0.01*np.sin(2*X[0])*np.cos(X[1]/2)
    c.Ur[2] = 1.0 - 0.5*np.tanh((X[1]-2*np.pi)/0.5)
    c.Ur_hat[:] = c.VM.forward(c.Ur)
    u0 = np.array(c.Ur_hat)
    config.params.t, config.params.tstep = 0.0, 0
    solve(solver, c)
    ref = o.solve(u0, 'Bq2D', 10, 0.005, 0.01, 0.2, 0.7)
    assert rel_l2(np.array(c.Ur_hat), ref) < 1e-11
    rho = solver.get_rho(**c)
    assert rel_l2(rho, o.backward(ref[2])) < 1e-11
