"""Velocity-vorticity solver module on the B200 path (reference solvers/VV.py: get_context :23-38,
compute_velocity :52-67, get_velocity/get_divergence/get_curl :69-83, getConvection :85-103,
add_linear :105-110, ComputeRHS :112-146).  The primary variable is the vorticity W_hat."""
from shenfun import VectorSpace, Array, Function
from .spectralinit import *          # noqa: F401,F403
from . import _common
from ._common import device_state    # noqa: F401
from .NS import end_of_tstep, set_velocity, compute_curl, NSFile   # noqa: F401  (VV builds on NS, reference VV.py:19)

_last_context = None


def get_context():
    """NS context without pressure, plus the vorticity W_hat as primary variable (VV.py:23-38)."""
    global _last_context
    float, complex, mpitype = datatypes(params.precision)
    collapse_fourier = params.dealias != '3/2-rule'
    dim = len(params.N)
    V, T, Tp, _engine = _common.build_spaces(comm, params, float, 'VV')
    VT = VectorSpace(T)
    VTp = VectorSpace(Tp)
    mask = T.get_mask_nyquist() if params.mask_nyquist else None
    X, K, K2, K_over_K2 = _common.wavenumber_arrays(T, VT, float)
    U = Array(VT)
    U_hat = Function(VT)
    u_dealias = Array(VTp)
    dU = Function(VT)
    curl = Array(VT)
    Source = Function(VT)
    work = work_arrays()
    W_hat = Function(VT, buffer=_common.pinned_like(VT.shape(True), complex)[0])
    u = W_hat                         # primary variable
    hdf5file = VVFile(config.params.solver,
                      checkpoint={'space': VT, 'data': {'0': {'curl': [W_hat]}}},
                      results={'space': VT, 'data': {'U': [U], 'curl': [curl]}})
    c = config.AttributeDict(locals())
    _last_context = c
    device_state(c)
    return c


class VVFile(HDF5File):
    def update_components(self, **context):
        get_velocity(**context)
        get_curl(**context)


def compute_velocity(U, w_hat, work, VT, K_over_K2):
    """u_hat = 1j*(k x w_hat)/k^2, u = F^-1(u_hat)"""
    v_hat = work[(w_hat, 1, True)]
    v_hat = cross2(v_hat, K_over_K2, w_hat)
    return VT.backward(v_hat, U)


def get_velocity(W_hat, U, work, VT, K_over_K2, **context):
    return compute_velocity(U, W_hat, work, VT, K_over_K2)


def get_divergence(T, K, U_hat, W_hat, **context):
    div_u = Array(T)
    U_hat = cross2(U_hat, K, W_hat)
    return T.backward(1j*(K[0]*U_hat[0]+K[1]*U_hat[1]+K[2]*U_hat[2]), div_u)


def get_curl(curl, W_hat, VT, **context):
    return VT.backward(W_hat, curl)


def getConvection(convection):
    if convection in ('Standard', 'Divergence', 'Skewed'):
        raise NotImplementedError
    return _common.Convection(convection)


def add_linear(rhs, w_hat, nu, K2, Source):
    """Fused into the last transform pass of ComputeRHS (kernel family vv_f0)."""
    raise NotImplementedError('add_linear is fused into ComputeRHS on the B200 path')


add_linear._sdns_builtin = True


def ComputeRHS(rhs, w_hat, solver, work, Tp, VT, VTp, K, K2, K_over_K2,
               Source, u_dealias, mask, **context):
    """rhs = 1j*K x F(u x w) masked, minus nu k^2 w_hat, plus Source."""
    if not getattr(getattr(solver, 'conv', None), '_sdns_builtin', True) or \
            not getattr(solver.add_linear, '_sdns_builtin', False):
        raise NotImplementedError('overriding conv/add_linear is not supported by the fused CUDA RHS')
    return _common.run_rhs(_common.dev_of(context), rhs, w_hat, Source, None)
