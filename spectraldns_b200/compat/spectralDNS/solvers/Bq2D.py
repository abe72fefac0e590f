"""2-D Navier-Stokes solver with the Boussinesq model on the B200 path.

Same module-level interface as the reference's solvers/Bq2D.py (get_context :13-99, get_Ur / get_rho / get_velocity
:101-119, getConvection :121-137, add_pressure_diffusion :139-156, ComputeRHS :158-186).  The state is
(u0, u1, rho); ComputeRHS -- u x curl and the density flux on the dealiased space, Nyquist mask, pressure with the
buoyancy term, diffusion of velocity and density -- is five CUDA launches behind sdns2d_compute_rhs."""
from shenfun import CompositeSpace
from .NS2D import *                   # noqa: F401,F403
from . import _common
from ._common import device_state     # noqa: F401

_last_context = None


def get_context():
    """Spaces, wavenumbers and solution arrays of the Bq2D solver, as an attribute dict."""
    global _last_context
    float, complex, mpitype = datatypes(params.precision)
    dim = len(params.N)
    V, T, Tp, _engine = _common.build_spaces(comm, params, float, 'Bq2D')
    VT = VectorSpace(T)
    VM = CompositeSpace([T]*(dim+1))
    mask = T.get_mask_nyquist() if params.mask_nyquist else None
    VTp = VectorSpace(Tp)
    VMp = CompositeSpace([Tp]*(dim+1))
    X, K, K2, K_over_K2 = _common.wavenumber_arrays(T, VT, float)

    Ur = Array(VM)
    Ur_hat = Function(VM, buffer=_common.pinned_like(VM.shape(True), complex)[0])
    P = Array(T)
    P_hat = Function(T)
    curl = Array(T)
    W_hat = Function(T)
    ur_dealias = Array(VMp)
    # views into the large arrays
    rho = Ur[2]
    rho_hat = Ur_hat[2]
    U = Ur[:2]
    U_hat = Ur_hat[:2]
    u = Ur_hat                        # primary variable
    dU = Function(VM)
    work = work_arrays()
    hdf5file = BqFile(config.params.solver,
                      checkpoint={'space': VM, 'data': {'0': {'Ur': [Ur_hat]}}},
                      results={'space': VM, 'data': {'UR': [Ur]}})
    context = config.AttributeDict(locals())
    context.pop('context', None)
    _last_context = context
    device_state(context)
    return context


class BqFile(HDF5File):
    """Transforms the stored components to physical space before a results write."""
    def update_components(self, Ur, Ur_hat, **context):
        Ur = Ur_hat.backward(Ur)


def get_Ur(Ur, Ur_hat, **context):
    return Ur_hat.backward(Ur)


def get_rho(Ur, Ur_hat, **context):
    Ur[2] = Ur_hat[2].backward(Ur[2])
    return Ur[2]


def get_velocity(Ur, Ur_hat, **context):
    Ur[0] = Ur_hat[0].backward(Ur[0])
    Ur[1] = Ur_hat[1].backward(Ur[1])
    return Ur[:2]


def add_pressure_diffusion(rhs, ur_hat, P_hat, K_over_K2, K, K2, nu, Ri, Pr):
    """Pressure with the buoyancy term, diffusion of velocity and density (reference Bq2D.py:139-156,
    cython_solvers.in:82-103).  ComputeRHS fuses it; called on its own it is one launch behind
    sdns2d_add_pressure_diffusion on host arrays staged in and out."""
    from spectralDNS.maths import _engine_for
    eng = _engine_for(ur_hat if hasattr(ur_hat, '_space') else rhs)
    p = eng.plan
    p.use_current_stream()
    p.Ri, p.Pr = float(Ri), float(Pr)
    d_r = eng.upload('apd_rhs', rhs, p.complex, p.tcomplex)
    d_u = eng.upload('apd_u', ur_hat, p.complex, p.tcomplex)
    d_p = eng.stage('apd_p', p.spectral_shape, p.tcomplex)
    p.add_pressure_diffusion(d_r, d_u, float(nu), d_p)
    rhs[...] = d_r.cpu().numpy()
    P_hat[...] = d_p.cpu().numpy()
    return rhs


add_pressure_diffusion._sdns_builtin = True


def ComputeRHS(rhs, ur_hat, solver, work, K, K2, K_over_K2, P_hat, T, Tp,
               VM, VMp, ur_dealias, mask, **context):
    """rhs of the Boussinesq equations: convection of momentum and density, masked, pressure + buoyancy, diffusion."""
    if not getattr(getattr(solver, 'conv', None), '_sdns_builtin', True) or \
            not getattr(solver.add_pressure_diffusion, '_sdns_builtin', False):
        raise NotImplementedError('overriding conv/add_pressure_diffusion is not supported by the fused CUDA RHS')
    return _common.run_rhs(_common.dev_of(context), rhs, ur_hat, None, P_hat)
