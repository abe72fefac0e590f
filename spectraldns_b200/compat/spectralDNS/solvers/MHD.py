"""Magnetohydrodynamics solver module (Elsasser / divergence form) on the B200 path (reference
solvers/MHD.py: get_context :13-77, set_Elsasser :89-97, divergenceConvection :99-110,
getConvection :112-130, add_pressure_diffusion :132-149, ComputeRHS :151-176).  State: UB_hat with
six components (velocity, magnetic field)."""
from shenfun import VectorSpace, Array, Function, CompositeSpace
from .spectralinit import *          # noqa: F401,F403
from . import _common
from ._common import device_state    # noqa: F401
from .NS import end_of_tstep         # noqa: F401

_last_context = None


def get_context():
    global _last_context
    float, complex, mpitype = datatypes(params.precision)
    collapse_fourier = params.dealias != '3/2-rule'
    dim = len(params.N)
    V, T, Tp, _engine = _common.build_spaces(comm, params, float, 'MHD')
    VT = VectorSpace(T)
    VM = CompositeSpace([T]*2*dim)
    mask = T.get_mask_nyquist() if params.mask_nyquist else None
    VTp = VectorSpace(Tp)
    VMp = CompositeSpace([Tp]*2*dim)
    X, K, K2, K_over_K2 = _common.wavenumber_arrays(T, VT, float)

    UB = Array(VM)
    P = Array(T)
    curl = Array(VT)
    UB_hat = Function(VM, buffer=_common.pinned_like(VM.shape(True), complex)[0])
    P_hat = Function(T)
    dU = Function(VM)
    Source = Array(VM)
    ub_dealias = Array(VMp)
    ZZ_hat = np.zeros((3, 3) + Tp.shape(True), dtype=complex)
    U, U_hat = UB[:3], UB_hat[:3]
    B, B_hat = UB[3:], UB_hat[3:]
    u = UB_hat                        # primary variable
    hdf5file = MHDFile(config.params.solver,
                       checkpoint={'space': VM, 'data': {'0': {'UB': [UB_hat]}}},
                       results={'space': VM, 'data': {'UB': [UB]}})
    context = config.AttributeDict(locals())
    context.pop('context', None)
    _last_context = context
    device_state(context)
    return context


class MHDFile(HDF5File):
    def update_components(self, UB, UB_hat, **kw):
        UB = UB_hat.backward(UB)


def get_divergence(T, K, U_hat, **context):
    div_u = Array(T)
    return T.backward(1j*(K[0]*U_hat[0]+K[1]*U_hat[1]+K[2]*U_hat[2]), div_u)


# Building blocks of the reference's MHD convection (MHD.py:89-110); the fused CUDA right-hand side (kernel families
# z_mhd / mhd_f0) does not call them.  Host arrays, any space with forward().
def set_Elsasser(c, ZZ, K):
    """c[:3] = -i/2 K_j (ZZ_ij + ZZ_ji),  c[3:] = i/2 K_j (ZZ_ji - ZZ_ij)   (MHD.py:89-97)"""
    for i in range(3):
        sym = sum(K[j]*(ZZ[i, j] + ZZ[j, i]) for j in range(3))
        asym = sum(K[j]*(ZZ[j, i] - ZZ[i, j]) for j in range(3))
        c[i] = -0.5j*sym
        c[3 + i] = 0.5j*asym
    return c


def divergenceConvection(c, z0, z1, Tp, K, ZZ_hat):
    """Elsasser products z0 = u + b, z1 = u - b: ZZ_hat[i, j] = forward(z0_i z1_j), then set_Elsasser   (MHD.py:99-110)"""
    for i in range(3):
        for j in range(3):
            ZZ_hat[i, j] = Tp.forward(z0[i]*z1[j], ZZ_hat[i, j])
    return set_Elsasser(c, ZZ_hat, K)


def getConvection(convection):
    if convection in ('Standard', 'Vortex', 'Skewed'):
        raise NotImplementedError
    return _common.Convection(convection)


def add_pressure_diffusion(rhs, ub_hat, nu, eta, K2, K, P_hat, K_over_K2):
    """Fused into the last transform pass of ComputeRHS (kernel family mhd_f0)."""
    raise NotImplementedError('add_pressure_diffusion is fused into ComputeRHS on the B200 path')


add_pressure_diffusion._sdns_builtin = True


def ComputeRHS(rhs, ub_hat, solver, Tp, VMp, K, K2, K_over_K2, P_hat,
               ub_dealias, ZZ_hat, mask, **context):
    """Elsasser products z+_i z-_j on the dealiased space, combined with i*K, masked, projected,
    minus nu k^2 u_hat and eta k^2 b_hat."""
    if not getattr(getattr(solver, 'conv', None), '_sdns_builtin', True) or \
            not getattr(solver.add_pressure_diffusion, '_sdns_builtin', False):
        raise NotImplementedError('overriding conv/add_pressure_diffusion is not supported by the fused CUDA RHS')
    return _common.run_rhs(_common.dev_of(context), rhs, ub_hat, None, P_hat)
