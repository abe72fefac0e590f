"""Classical Navier-Stokes solver module on the B200 path.

Same module-level interface as the reference's solvers/NS.py (get_context :12-72, get_curl /
get_velocity / get_pressure / set_velocity / get_divergence :86-110, end_of_tstep :112-122,
getConvection :164-201, add_pressure_diffusion :203-217, ComputeRHS :219-261).  The arithmetic of
ComputeRHS -- u x curl(u) on the dealiased space, Nyquist mask, pressure projection, viscous term,
Source -- is five CUDA launches behind sdns_compute_rhs; the context's numpy arrays are host
mirrors of device-resident state."""
from shenfun import VectorSpace, Array, Function
from .spectralinit import *          # noqa: F401,F403
from . import _common
from ._common import device_state    # noqa: F401  (solve() and getintegrator() use solver.device_state)

_last_context = None


def get_context():
    """Spaces, wavenumbers and solution arrays of the NS solver, as an attribute dict."""
    global _last_context
    float, complex, mpitype = datatypes(params.precision)
    collapse_fourier = params.dealias != '3/2-rule'
    dim = len(params.N)
    V, T, Tp, _engine = _common.build_spaces(comm, params, float, 'NS')
    VT = VectorSpace(T)
    VTp = VectorSpace(Tp)
    mask = T.get_mask_nyquist() if params.mask_nyquist else None
    X, K, K2, K_over_K2 = _common.wavenumber_arrays(T, VT, float)

    U = Array(VT)
    U_hat = Function(VT, buffer=_common.pinned_like(VT.shape(True), complex)[0])
    P = Array(T)
    P_hat = Function(T)
    u_dealias = Array(VTp)
    u = U_hat                         # primary variable
    dU = Function(VT)                 # right hand side
    curl = Array(VT)
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


class NSFile(HDF5File):
    """Transforms the stored components to physical space before a results write."""
    def update_components(self, **context):
        get_velocity(**context)
        get_pressure(**context)


def get_curl(curl, U_hat, work, VT, K, **context):
    return compute_curl(curl, U_hat, work, VT, K)


def get_velocity(U, U_hat, VT, **context):
    return VT.backward(U_hat, U)


def get_pressure(P, P_hat, T, **context):
    return T.backward(-1j*P_hat, P)


def set_velocity(U, U_hat, VT, **context):
    return VT.forward(U, U_hat)


def get_divergence(T, K, U_hat, mask, **context):
    div_u = Array(T)
    return T.backward(1j*(K[0]*U_hat[0]+K[1]*U_hat[1]+K[2]*U_hat[2]), div_u)


def end_of_tstep(context):
    """Shorten the last step so that the run ends on params.T (used by adaptive runs)."""
    if abs(params.t - params.T) < 1e-12:
        return True
    if (abs(params.t + params.dt - params.T) < 1e-12 or params.t + params.dt >= params.T + 1e-12):
        params.dt = params.T - params.t
    return False


def compute_curl(c, a, work, T, K):
    """c = F^-1(1j*K x a)"""
    curl_hat = work[(a, 0, False)]
    curl_hat = cross2(curl_hat, K, a)
    return T.backward(curl_hat, c)


# The building blocks getConvection's closures are made of in the reference (NS.py:131-162).  The fused CUDA right-hand
# side does not call them; they are here for user code that does, and work on host arrays with whatever space is
# passed in (its forward / backward are the B200 transforms for the spaces of get_context()).
def Cross(c, a, b, work, T):
    """c = T.forward(a x b)   (NS.py:131-136)"""
    prod = np.empty_like(a)
    prod[0] = a[1]*b[2] - a[2]*b[1]
    prod[1] = a[2]*b[0] - a[0]*b[2]
    prod[2] = a[0]*b[1] - a[1]*b[0]
    return T.forward(prod, c)


def standard_convection(rhs, u_dealias, U_hat, work, Tp, K):
    """rhs_i = forward(u_j d u_i / d x_j)   (NS.py:138-145)"""
    grad = np.empty_like(u_dealias[0])
    for i in range(3):
        acc = np.zeros_like(u_dealias[0])
        for j in range(3):
            grad = Tp.backward(1j*K[j]*U_hat[i], grad)
            acc += u_dealias[j]*grad
        rhs[i] = Tp.forward(acc, rhs[i])
    return rhs


def divergence_convection(rhs, u_dealias, work, Tp, K, add=False):
    """rhs_i (+)= i K_j forward(u_i u_j), six transforms for the symmetric product   (NS.py:147-162)"""
    if not add:
        rhs.fill(0)
    uu = np.empty_like(rhs[0])
    for i in range(3):
        for j in range(i, 3):
            uu = Tp.forward(u_dealias[i]*u_dealias[j], uu)
            rhs[i] += 1j*K[j]*uu
            if j != i:
                rhs[j] += 1j*K[i]*uu
    return rhs


def getConvection(convection):
    """Nonlinear term selector: 'Vortex' u x curl(u) (default), 'Standard' u_j du_i/dx_j,
    'Divergence' d(u_i u_j)/dx_j, 'Skewed' their mean -- all compiled into the CUDA pipeline
    (kernel families ns_b0/z_cross, ns_grad_b0/z_dot, z_uu/nsdiv_f0)."""
    if convection not in ('Vortex', 'Standard', 'Divergence', 'Skewed'):
        raise NotImplementedError(convection)
    return _common.Convection(convection)


def add_pressure_diffusion(rhs, u_hat, nu, K2, K, P_hat, K_over_K2):
    """rhs -= P_hat*K + nu*K2*u_hat with P_hat = sum(rhs*K_over_K2, 0) (reference NS.py:203-217,
    cython_solvers.in:40-80).  ComputeRHS fuses this into its last transform pass (kernel family ns_f0);
    called on its own -- the fine-grained plug-in surface of optimization/__init__.py:12-55 -- it runs as
    one launch behind sdns_add_pressure_diffusion on host arrays staged in and out."""
    from spectralDNS.maths import _engine_for
    eng = _engine_for(u_hat if hasattr(u_hat, '_space') else rhs)
    p = eng.plan
    p.use_current_stream()
    d_r = eng.upload('apd_rhs', rhs, p.complex, p.tcomplex)
    d_u = eng.upload('apd_u', u_hat, p.complex, p.tcomplex)
    d_p = eng.stage('apd_p', p.spectral_shape, p.tcomplex)
    p.add_pressure_diffusion(d_r, d_u, float(nu), d_p)
    rhs[...] = d_r.cpu().numpy()
    P_hat[...] = d_p.cpu().numpy()
    return rhs


add_pressure_diffusion._sdns_builtin = True


def ComputeRHS(rhs, u_hat, solver, work, Tp, VTp, P_hat, K, K2, u_dealias,
               K_over_K2, Source, mask, **context):
    """rhs = F(u x curl u) masked, minus pressure gradient and nu k^2 u_hat, plus Source."""
    if not getattr(getattr(solver, 'conv', None), '_sdns_builtin', True) or \
            not getattr(solver.add_pressure_diffusion, '_sdns_builtin', False):
        raise NotImplementedError('overriding conv/add_pressure_diffusion is not supported by the fused CUDA RHS')
    return _common.run_rhs(_common.dev_of(context), rhs, u_hat, Source, P_hat)
