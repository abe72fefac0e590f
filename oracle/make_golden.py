"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

It puts oracle/shim (numpy stand-ins for shenfun / mpi4py / mpi4py_fft, which are absent from
the image) and /root/reference on sys.path, imports the reference's own spectralDNS package
(solvers/NS.py, VV.py, MHD.py, maths/integrators.py, config.py, __init__.py -- none of them
copied or modified), runs each case through the reference's get_solver / get_context /
ComputeRHS / solve, and stores inputs + outputs as small fixtures.  It also cross-checks the
restatement in oracle/sdns_oracle.py against the reference's output and prints the relative L2
difference for each case.
"""
import os
import sys
import io
import contextlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path[:0] = [os.path.join(HERE, 'shim'), REF, os.path.join(REF, 'tests'), HERE]

from spectralDNS import config, get_solver, solve   # noqa: E402  (the reference's package)
import sdns_oracle as so                            # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def rel(a, b):
    return float(np.linalg.norm((a-b).ravel())/max(np.linalg.norm(b.ravel()), 1e-300))


def ref_solver(name, args):
    config.update({'nu': 0.000625, 'dt': 0.01, 'T': 0.1,
                   'convection': 'Divergence' if name == 'MHD' else 'Vortex'})
    with contextlib.redirect_stdout(io.StringIO()):
        solver = get_solver(parse_args=list(args)+[name])
        ctx = solver.get_context()
    config.params.t = 0.0
    config.params.tstep = 0
    return solver, ctx


def run_ref(solver, ctx, nsteps):
    config.params.t = 0.0
    config.params.tstep = 0
    config.params.T = config.params.dt*nsteps
    with contextlib.redirect_stdout(io.StringIO()):
        solve(solver, ctx)


def oracle_for(ctx, p):
    return so.Oracle(p.N, p.L, p.precision, p.dealias, p.mask_nyquist)


def case_tg(name, margs, fname, nsteps=10):
    """Taylor-Green regression of tests/test_NSVV.py:36-55 / tests/TG.py:116-126."""
    solver, c = ref_solver(name, margs)
    p = config.params
    o = oracle_for(c, p)
    U0 = so.taylor_green(o)
    c.U[:] = U0
    solver.set_velocity(**c)
    if name == 'VV':
        solver.cross2(c.W_hat, c.K, c.U_hat)
    u0_hat = np.array(c.u)
    run_ref(solver, c, nsteps)
    u_hat = np.array(c.u)
    U = np.array(solver.get_velocity(**c))
    curl = np.array(solver.get_curl(**c))
    k = float(np.sum(U.astype(np.float64)**2)/np.prod(p.N)/2)
    w = float(np.sum(curl.astype(np.float64)**2)/np.prod(p.N)/2)
    ntol = 7 if p.precision == 'double' else 5         # the reference only tests double
    assert round(w - 0.375249930801, ntol) == 0, w    # tests/TG.py:125
    assert round(k - 0.124953117517, ntol) == 0, k    # tests/TG.py:126
    mine = o.solve(u0_hat, name, nsteps, p.dt, p.nu)
    print('%-28s k=%.12f w=%.12f  oracle-vs-reference rel L2 = %.2e' % (fname, k, w, rel(mine, u_hat)))
    np.savez_compressed(os.path.join(OUT, fname), solver=name, N=p.N, L=p.L, precision=p.precision,
                        dealias=p.dealias, nu=float(p.nu), dt=float(p.dt), nsteps=nsteps,
                        u_hat=u_hat, k=k, w=w)    # u0_hat = forward(taylor_green), regenerated in tests


def case_integrator(integrator, fname, margs):
    """tests/test_NSVV.py:73-92: the same TG problem under every explicit integrator (run to T=0.1;
    the adaptive one chooses its own steps)."""
    solver, c = ref_solver('NS', margs + ['--integrator', integrator])
    p = config.params
    o = oracle_for(c, p)
    f0 = so.isotropic_field(o, seed=5)
    c.u[:] = f0
    p.t, p.tstep, p.T, p.dt = 0.0, 0, 0.05, 0.005
    with contextlib.redirect_stdout(io.StringIO()):
        solve(solver, c)
    print('%-28s integrator %-13s steps=%d t=%.6f' % (fname, integrator, p.tstep, p.t))
    np.savez_compressed(os.path.join(OUT, fname), solver='NS', N=p.N, L=p.L, precision=p.precision,
                        dealias=p.dealias, nu=float(p.nu), dt=0.005, T=0.05, integrator=integrator,
                        nsteps=int(p.tstep), t_end=float(p.t), u0_hat=f0, u_hat=np.array(c.u))
    p.dt, p.T, p.integrator = 0.01, 0.1, 'RK4'


def case_mhd(margs, fname, nsteps=10):
    """tests/test_MHD.py:32-50 / tests/TGMHD.py:4-26."""
    solver, c = ref_solver('MHD', margs)
    p = config.params
    o = oracle_for(c, p)
    c.UB[:] = so.taylor_green_mhd(o)
    c.UB.forward(c.UB_hat)
    u0_hat = np.array(c.UB_hat)
    run_ref(solver, c, nsteps)
    u_hat = np.array(c.UB_hat)
    UB = np.array(c.UB_hat.backward(c.UB))
    k = float(np.sum(UB[:3].astype(np.float64)**2)/np.prod(p.N)/2)
    b = float(np.sum(UB[3:].astype(np.float64)**2)/np.prod(p.N)/2)
    assert round(k - 0.124565408177, 7) == 0, k       # tests/TGMHD.py:25
    assert round(b - 0.124637762143, 7) == 0, b       # tests/TGMHD.py:26
    mine = o.solve(u0_hat, 'MHD', nsteps, p.dt, p.nu, eta=p.eta)
    print('%-28s k=%.12f b=%.12f  oracle-vs-reference rel L2 = %.2e' % (fname, k, b, rel(mine, u_hat)))
    np.savez_compressed(os.path.join(OUT, fname), solver='MHD', N=p.N, L=p.L, precision=p.precision,
                        dealias=p.dealias, nu=float(p.nu), eta=float(p.eta), dt=float(p.dt),
                        nsteps=nsteps, u_hat=u_hat, k=k, b=b)


def case_broadband(name, margs, fname, nsteps=3, seed=0, convections=('Vortex',)):
    """Seeded broadband (demo/Isotropic.py:29-76 spectrum) field: one ComputeRHS per convection
    form plus nsteps RK4 steps.  TG cannot discriminate dealiasing conventions; this does."""
    solver, c = ref_solver(name, margs)
    p = config.params
    o = oracle_for(c, p)
    ncomp = 6 if name == 'MHD' else 3
    f0 = so.isotropic_field(o, seed=seed, ncomp=ncomp)
    if name == 'MHD':
        f0[3:] = so.isotropic_field(o, seed=seed+1, ncomp=3)*0.5
    if name == 'VV':
        f0 = o.cross2(o.K, f0)     # state is the vorticity
    c.u[:] = f0
    u0_hat = np.array(c.u)
    store = dict(solver=name, N=p.N, L=p.L, precision=p.precision, dealias=p.dealias,
                 nu=float(p.nu), dt=float(p.dt), nsteps=nsteps, u0_hat=u0_hat)
    if name == 'MHD':
        store['eta'] = float(p.eta)
    msg = []
    for conv in convections:
        p.convection = conv
        solver.conv = solver.getConvection(conv)
        rhs = np.array(solver.ComputeRHS(c.dU, c.u, solver, **c))
        store['rhs_'+conv] = rhs
        if name == 'NS':
            mine, P = o.ns_rhs(u0_hat, p.nu, conv, return_p=True)
            store['P_hat'] = np.array(c.P_hat)
        elif name == 'VV':
            mine = o.vv_rhs(u0_hat, p.nu)
        else:
            mine = o.mhd_rhs(u0_hat, p.nu, p.eta)
        msg.append('%s %.1e' % (conv, rel(mine, rhs)))
    p.convection = convections[0]
    c.u[:] = u0_hat
    run_ref(solver, c, nsteps)
    u_hat = np.array(c.u)
    store['u_hat'] = u_hat
    mine = o.solve(u0_hat, name, nsteps, p.dt, p.nu, eta=p.get('eta', None), convection=convections[0])
    print('%-28s rhs[%s]  %d steps rel L2 = %.2e' % (fname, ', '.join(msg), nsteps, rel(mine, u_hat)))
    np.savez_compressed(os.path.join(OUT, fname), **store)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    uni = ['--M', '4', '4', '4', '--L', '2*pi', '2*pi', '2*pi']
    non = ['--M', '6', '5', '4', '--L', '6*pi', '4*pi', '2*pi']
    m5 = ['--M', '5', '5', '5']
    # the reference's own regression matrix (tests/test_NSVV.py:27-30, tests/test_MHD.py:22-27)
    case_tg('NS', uni, 'tg_ns_16_double')
    case_tg('VV', uni, 'tg_vv_16_double')
    case_tg('NS', non, 'tg_ns_64x32x16_double', nsteps=10)
    case_tg('VV', non, 'tg_vv_64x32x16_double', nsteps=10)
    case_tg('NS', m5, 'tg_ns_32_double')                       # BASELINE.json configs[0]
    case_tg('NS', m5+['--precision', 'single'], 'tg_ns_32_single')
    case_tg('NS', m5+['--dealias', '3/2-rule'], 'tg_ns_32_double_pad')
    case_mhd(uni, 'tg_mhd_16_double')
    case_mhd(non, 'tg_mhd_64x32x16_double', nsteps=10)
    # broadband fields: field-level parity incl. dealiasing convention (small grids keep the
    # fixtures small; the spectra are dense so they do not compress)
    m4 = uni
    nonb = ['--M', '5', '4', '3', '--L', '6*pi', '4*pi', '2*pi']
    case_broadband('NS', m5, 'iso_ns_32_double', convections=('Vortex',))
    case_broadband('NS', m4, 'iso_ns_16_double', convections=('Vortex', 'Standard', 'Divergence', 'Skewed'))
    case_broadband('NS', m4+['--precision', 'single'], 'iso_ns_16_single')
    case_broadband('NS', m4+['--dealias', '3/2-rule'], 'iso_ns_16_double_pad')
    case_broadband('NS', m4+['--dealias', '3/2-rule', '--precision', 'single'], 'iso_ns_16_single_pad')
    case_broadband('NS', nonb, 'iso_ns_32x16x8_double')
    case_broadband('NS', m4+['--dealias', 'None', '--no-mask_nyquist'], 'iso_ns_16_double_nodealias')
    case_broadband('VV', m4, 'iso_vv_16_double')
    case_broadband('VV', m4+['--dealias', '3/2-rule'], 'iso_vv_16_double_pad')
    case_broadband('MHD', m4, 'iso_mhd_16_double', convections=('Divergence',))
    case_broadband('MHD', m4+['--dealias', '3/2-rule'], 'iso_mhd_16_double_pad', convections=('Divergence',))
    case_broadband('MHD', m4+['--precision', 'single'], 'iso_mhd_16_single', convections=('Divergence',))
    for integ in ('ForwardEuler', 'AB2', 'BS5_fixed', 'BS5_adaptive'):
        case_integrator(integ, 'integ_ns_16_%s' % integ.lower(), m4)
