"""2-D Navier-Stokes solver module on the B200 path.

Same module-level interface as the reference's solvers/NS2D.py (get_context :13-18, get_curl :20-24,
get_divergence :26-30, getConvection :32-51) on top of what it inherits from solvers/NS.py (get_velocity,
get_pressure, set_velocity, add_pressure_diffusion, ComputeRHS).  The curl of a 2-D field is a scalar; ComputeRHS --
u x curl on the dealiased space, Nyquist mask, pressure projection, viscous term, Source -- is five CUDA launches
behind sdns2d_compute_rhs (csrc/sdns2d_api.cu)."""
from .NS import *                     # noqa: F401,F403
from . import _common
from ._common import device_state     # noqa: F401

_last_context = None


def get_context():
    """Spaces, wavenumbers and solution arrays of the NS2D solver, as an attribute dict."""
    global _last_context
    float, complex, mpitype = datatypes(params.precision)
    dim = len(params.N)
    V, T, Tp, _engine = _common.build_spaces(comm, params, float, 'NS2D')
    VT = VectorSpace(T)
    VTp = VectorSpace(Tp)
    mask = T.get_mask_nyquist() if params.mask_nyquist else None
    X, K, K2, K_over_K2 = _common.wavenumber_arrays(T, VT, float)

    U = Array(VT)
    U_hat = Function(VT, buffer=_common.pinned_like(VT.shape(True), complex)[0])
    P = Array(T)
    P_hat = Function(T)
    u_dealias = Array(VTp)
    u = U_hat
    dU = Function(VT)
    curl = Array(T)
    W_hat = Function(T)
    Source = Function(VT)
    work = work_arrays()
    hdf5file = NSFile(config.params.solver,
                      checkpoint={'space': VT, 'data': {'0': {'U': [U_hat]}}},
                      results={'space': VT, 'data': {'U': [U], 'P': [P]}})
    context = config.AttributeDict(locals())
    context.pop('context', None)
    _last_context = context
    device_state(context)
    return context


def get_curl(curl, W_hat, U_hat, work, T, K, **context):
    W_hat[:] = 0
    W_hat = cross2(W_hat, K, U_hat)
    curl = W_hat.backward(curl)
    return curl


def get_divergence(T, K, U_hat, mask, **context):
    div_u = Array(T)
    return T.backward(1j*(K[0]*U_hat[0]+K[1]*U_hat[1]), div_u)


def getConvection(convection):
    """Only 'Vortex' exists in two dimensions (reference NS2D.py:34-35); it is compiled into the CUDA pipeline
    (kernel family z_ns2d)."""
    if convection != 'Vortex':
        raise NotImplementedError(convection)
    return _common.Convection(convection)
