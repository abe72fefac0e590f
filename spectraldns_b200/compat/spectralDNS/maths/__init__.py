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
    if len(p.spectral_shape) == 2:
        # 2-D: scalar c = 1j*(a0 b1 - a1 b0) (cython_maths.in:105-147).  With the wavenumber list K this is one launch
        # (sdns2d_cross2); a dense real `a` is user-side arithmetic and stays numpy.
        if isinstance(a, (list, tuple)):
            db = eng.upload('x2b', np.asarray(b)[:2], p.complex, p.tcomplex)
            dc = eng.stage('x2c', p.spectral_shape, p.tcomplex)
            p.cross2(dc, db)
            c[...] = dc.cpu().numpy()
        else:
            c[...] = 1j*(a[0]*np.asarray(b)[1] - a[1]*np.asarray(b)[0])
        return c
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
        if hasattr(plan, 'set_physics'):
            plan.set_physics(params)
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
    elif name in ('BS5_adaptive', 'BS5_fixed') and not hasattr(plan, 'lincomb'):
        raise NotImplementedError('%s is available for the triply periodic solvers; the 2-D solvers provide RK4, '
                                  'ForwardEuler and AB2' % name)
    elif name in ('BS5_adaptive', 'BS5_fixed'):
        integrate = _bs5(name == 'BS5_adaptive', rhs, u0, solver, context, dev, before, after, eta)
    else:
        raise NotImplementedError("unknown integrator %r" % name)
    integrate.__name__ = name
    return integrate


# Bogacki-Shampine 5(4) pair, first-same-as-last (the tableau the reference hard-codes at
# maths/integrators.py:199-208)
_BS5_A = [[0, 0, 0, 0, 0, 0, 0, 0],
          [1/6, 0, 0, 0, 0, 0, 0, 0],
          [2/27, 4/27, 0, 0, 0, 0, 0, 0],
          [183/1372, -162/343, 1053/1372, 0, 0, 0, 0, 0],
          [68/297, -4/11, 42/143, 1960/3861, 0, 0, 0, 0],
          [597/22528, 81/352, 63099/585728, 58653/366080, 4617/20480, 0, 0, 0],
          [174197/959244, -30942/79937, 8152137/19744439, 666106/1039181, -29421/29068, 482048/414219, 0, 0],
          [587/8064, 0, 4440339/15491840, 24353/124800, 387/44800, 2152/5985, 7267/94080, 0]]
_BS5_B = [587/8064, 0, 4440339/15491840, 24353/124800, 387/44800, 2152/5985, 7267/94080, 0]
_BS5_BHAT = [2479/34992, 0, 123/416, 612941/3411720, 43/1440, 2272/6561, 79937/1113912, 3293/556956]


def _bs5(adaptive, rhs, u0, solver, context, dev, before, after, eta):
    """adaptiveRK of the reference (maths/integrators.py:15-147) with every stage vector resident on
    the GPU: stage values through sdns_lincomb, right-hand sides through sdns_compute_rhs, the
    error estimate through sdns_errnorm.  Step-size controller, FSAL rotation (`offset`) and the
    rejected-step callback protocol are the reference's."""
    params = solver.params
    plan = dev.plan
    fl = context.float
    A = np.array(_BS5_A, dtype=fl)
    b = np.array(_BS5_B, dtype=fl)
    bhat = np.array(_BS5_BHAT, dtype=fl)
    s = A.shape[0]
    err_order = 4
    offset = [0]
    fY = [plan.empty_spectral(dev.ncomp) for _ in range(s)]
    ytmp = plan.empty_spectral(dev.ncomp)
    unew = plan.empty_spectral(dev.ncomp)
    err = plan.empty_spectral(dev.ncomp)
    default_cb = getattr(solver.additional_callback, '_sdns_default', False)
    ntot = float(np.prod(context.T.shape(True)))

    def callback():
        if not default_cb:
            dev.device_newer = True
            dev.sync_to_host()
            solver.additional_callback(context)

    def integrate():
        src = before()
        dt = float(params.dt)
        tstep = int(params.tstep)
        aTOL = rTOL = float(params.TOL)
        facmax, fac, facmin = 2, 0.8, 0.01
        nu = float(params.nu)
        while True:
            dt_prev = dt
            offset[0] = (offset[0] - 1) % s                       # first-same-as-last rotation
            for i in range(s):
                slot = fY[(i + offset[0]) % s]
                if tstep == 0 or i != 0:
                    terms = [(dt*float(A[i, j]), fY[(j + offset[0]) % s]) for j in range(i) if A[i, j] != 0]
                    plan.lincomb(ytmp, dev.u, [c for c, _ in terms], [x for _, x in terms])
                    plan.compute_rhs(slot, ytmp, nu, eta(), source=src)
                if i == 0:
                    if not default_cb:
                        context.fu0 = slot.cpu().numpy()
                    callback()
            plan.lincomb(unew, dev.u, [dt*float(b[j]) for j in range(s)], [fY[(j + offset[0]) % s] for j in range(s)])
            plan.lincomb(err, None, [dt*float(b[j] - bhat[j]) for j in range(s)],
                         [fY[(j + offset[0]) % s] for j in range(s)])
            nsq = plan.errnorm(dev.u, unew, err, aTOL, rTOL)
            nsq = np.array([solver.comm.allreduce(float(v)) for v in nsq])
            est = float(np.max(np.sqrt(nsq)))/np.sqrt(ntot)
            exponent = 1.0/(err_order + 1)
            factor = min(facmax, max(facmin, fac*pow((1/est), exponent))) if est > 0 else facmax
            if adaptive:
                dt = dt*factor
                if est > 1.0:
                    facmax = 1
                    context.is_step_rejected_callback = True
                    context.dt_rejected = dt_prev
                    callback()
                    offset[0] += 1
                    continue
            break
        dev.u.copy_(unew)
        after()
        return u0, fl(dt), fl(dt_prev)

    return integrate
