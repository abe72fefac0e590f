"""cross1, cross2, project and getintegrator with the reference's signatures
(maths/cross.py:16-35, maths/maths.py:8-11, maths/integrators.py:177-239).  Host arrays in and out;
the arithmetic runs on the GPU through the C ABI."""
import numpy as np

from spectralDNS import config

__all__ = ['cross1', 'cross2', 'project', 'getintegrator']


def _engine_for(arr):
    """The CUDA engine of the array's function space, else the one of the current parameters."""
    from spectraldns_b200.spaces import Engine, _scalar_space
    sp = getattr(arr, '_space', None)
    if sp is not None:
        return _scalar_space(sp).engine
    p = config.params
    return Engine.get(p.N, p.L, p.precision, p.dealias, p.solver, p.mask_nyquist, p.decomposition)


def cross1(c, a, b):
    """c = a x b for real (3, ...) arrays."""
    eng = _engine_for(c)
    p = eng.plan
    p.use_current_stream()
    da = eng.upload('x1a', a, p.float, p.tfloat)
    db = eng.upload('x1b', b, p.float, p.tfloat)
    dc = eng.stage('x1c', da.shape, p.tfloat)
    p.cross1(dc, da, db)
    c[...] = dc.cpu().numpy()
    return c


def cross2(c, a, b):
    """c = 1j*(a x b); a real -- the list K of broadcast wavenumber arrays or a dense (3, ...) array
    such as K_over_K2 -- and b complex."""
    eng = _engine_for(b if hasattr(b, '_space') else c)
    p = eng.plan
    p.use_current_stream()
    db = eng.upload('x2b', b, p.complex, p.tcomplex)
    dc = eng.stage('x2c', db.shape, p.tcomplex)
    if isinstance(a, (list, tuple)):
        dense = np.empty((3,)+tuple(db.shape[1:]), dtype=p.float)
        for i in range(3):
            dense[i] = a[i]
        a = dense
    da = eng.upload('x2a', a, p.float, p.tfloat)
    p.cross2_dense(dc, da, db)
    c[...] = dc.cpu().numpy()
    return c


def project(u, K, K_over_K2):
    """Project u onto the divergence-free space (in place)."""
    eng = _engine_for(u)
    p = eng.plan
    p.use_current_stream()
    du = eng.upload('prj', u, p.complex, p.tcomplex)
    p.project(du)
    u[...] = du.cpu().numpy()
    return u


def getintegrator(rhs, u0, solver, context):
    """Return the zero-argument integrate() of params.integrator.  integrate() returns
    (u0, dt, dt_took) like the reference; the stage updates run fused on the device."""
    params = solver.params
    name = params.integrator
    dev = solver.device_state(context)
    plan = dev.plan

    def eta():
        return float(params.eta) if 'eta' in params else 0.0

    def before():
        if dev.host_dirty or not dev.managed:
            dev.upload_state()
            dev.refresh_source(context.get('Source', None))
        plan.use_current_stream()
        return dev.source if dev.source_active else None

    def after():
        dev.device_newer = True
        if not dev.managed:
            dev.sync_to_host()

    if name == 'RK4':
        def integrate():
            src = before()
            plan.rk4_step(dev.u, dev.u1, dev.u2, float(params.dt), float(params.nu), eta(), src)
            after()
            return u0, params.dt, params.dt
    elif name == 'ForwardEuler':
        def integrate():
            src = before()
            plan.euler_step(dev.u, dev.rhs_buffer(), float(params.dt), float(params.nu), eta(), src)
            after()
            return u0, params.dt, params.dt
    elif name == 'AB2':
        def integrate():
            src = before()
            plan.ab2_step(dev.u, dev.u1, dev.rhs_buffer(), float(params.dt), int(params.tstep),
                          float(params.nu), eta(), src)
            after()
            return u0, params.dt, params.dt
    else:
        raise NotImplementedError("integrator %r is not on the B200 path yet (RK4, ForwardEuler, AB2 are)" % name)
    integrate.__name__ = name
    return integrate
