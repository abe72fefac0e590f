"""spectralDNS.maths.integrators: getintegrator (the device-resident integrators of the package) plus the explicit
one-step functions of the reference under their own names (maths/integrators.py:150-175), for user code that calls
them directly.  These work on host arrays through solver.ComputeRHS, exactly like the reference's Python versions;
the time loop of solve() does not use them -- it uses the fused CUDA stage updates behind getintegrator."""
from . import getintegrator          # noqa: F401

__all__ = ['getintegrator', 'RK4', 'ForwardEuler', 'AB2']


def RK4(u0, u1, u2, rhs, a, b, dt, solver, context):
    """Classical fourth-order Runge-Kutta step (maths/integrators.py:150-159): returns (u0, dt, dt)."""
    u1[...] = u0
    u2[...] = u0
    for stage in range(4):
        rhs = solver.ComputeRHS(rhs, u0, solver, **context)
        if stage < 3:
            u0[...] = u1 + (b[stage]*dt)*rhs
        u2 += (a[stage]*dt)*rhs
    u0[...] = u2
    return u0, dt, dt


def ForwardEuler(u0, rhs, dt, solver, context):
    """maths/integrators.py:161-165"""
    rhs = solver.ComputeRHS(rhs, u0, solver, **context)
    u0 += dt*rhs
    return u0, dt, dt


def AB2(u0, u1, rhs, dt, tstep, solver, context):
    """Second-order Adams-Bashforth, started with one Euler step; u1 keeps dt*rhs of the previous step
    (maths/integrators.py:167-175)."""
    rhs = solver.ComputeRHS(rhs, u0, solver, **context)
    step = dt*rhs
    if tstep == 0:
        u0 += step
    else:
        u0 += 1.5*step - 0.5*u1
    u1[...] = step
    return u0, dt, dt
