"""spectralDNS front end on the B200 path: get_solver() and solve() with the reference's
signatures and call order (reference spectralDNS/__init__.py:19-66 and :69-123), so demo/TG.py,
demo/TGMHD.py and demo/Isotropic.py run unchanged.  The time loop keeps the solution on the GPU;
host arrays in the context are refreshed only when a user callback or a file write needs them
(spectraldns_b200/device_state.py)."""
import importlib

from . import config

__version__ = '1.4.0+b200'

_SOLVERS = ('NS', 'VV', 'MHD', 'NS2D', 'Bq2D')


def get_solver(update=None, regression_test=None, additional_callback=None,
               mesh='triplyperiodic', parse_args=None):
    """Parse the command line (or `parse_args`) into config.params, import the solver module
    named by the positional sub-command and install the user callbacks on it."""
    if parse_args is not None and not isinstance(parse_args, list):
        raise AssertionError('parse_args must be a list of strings or None')
    namespace = getattr(config, mesh).parse_args(parse_args)
    config.params.update(vars(namespace))
    name = config.params.solver
    if name not in _SOLVERS:
        raise AttributeError("Wrong solver! The B200 path provides %s, got %r" % (', '.join(_SOLVERS), name))
    solver = importlib.import_module('spectralDNS.solvers.' + name)
    for attr, fn in (('update', update), ('regression_test', regression_test),
                     ('additional_callback', additional_callback)):
        if fn:
            setattr(solver, attr, fn)
    config.solver = solver
    config.mesh = mesh
    return solver


def solve(solver, context):
    """Integrate from params.t to params.T, calling the solver's hooks every step in the
    reference's order: integrate, update, hdf5file.update, timer, end_of_tstep."""
    params = solver.params
    solver.timer = solver.Timer()
    solver.conv = solver.getConvection(params.convection)
    integrate = solver.getintegrator(context.dU, context.u, solver, context)
    dev = solver.device_state(context)
    user_update = not getattr(solver.update, '_sdns_default', False)
    # SDNS_LAZY_STATE=1: the host mirror of the state is refreshed when host code touches it instead of before every
    # callback, and the forcing expressions of demo/Isotropic.py's update() run on the device (spectraldns_b200/spaces.py)
    from spectraldns_b200.spaces import lazy_state_enabled
    lazy = user_update and lazy_state_enabled()
    dev.begin_solve(lazy)
    dt_in = params.dt
    try:
        while params.t + params.dt <= params.T + 1e-12:
            u, params.dt, dt_took = integrate()
            params.t += dt_took
            params.tstep += 1
            if user_update and lazy:
                solver.update(context)
            elif user_update:
                dev.sync_to_host()
                dev.host_touched()          # the callback may change the host state and then call ComputeRHS / transforms on it
                solver.update(context)
                dev.host_touched()
            context.hdf5file.update(params, **context)
            solver.timer()
            if not solver.profiler.getstats() and params.make_profile:
                solver.profiler.enable()
            if solver.end_of_tstep(context):
                break
    finally:
        dev.end_solve()
    params.dt = dt_in
    solver.timer.final(params.verbose)
    if params.make_profile:
        solver.results = solver.create_profile(solver.profiler)
    solver.regression_test(context)
    context.hdf5file.close()
