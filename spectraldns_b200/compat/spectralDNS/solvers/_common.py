"""Shared pieces of the NS / VV / MHD solver modules on the B200 path."""
import numpy as np

from shenfun import FunctionSpace, TensorProductSpace
from spectralDNS import config
from spectraldns_b200.spaces import Engine
from spectraldns_b200.device_state import DeviceState


def build_spaces(comm, params, float_, solver_name):
    """T and its dealiased companion Tp sharing one CUDA plan (reference solvers/NS.py:14-32)."""
    dim = len(params.N)
    assert dim in (2, 3), 'the B200 path covers the triply and the doubly periodic solvers'
    V = [FunctionSpace(params.N[i], 'F', domain=(0, params.L[i]),
                       dtype=(float_ if i == dim-1 else (np.complex64 if float_ == np.float32 else np.complex128)))
         for i in range(dim)]
    conv = params.convection if solver_name == 'NS' else None      # VV: Vortex, MHD: Divergence only
    if params.decomposition == 'pencil' and comm.Get_size() > 1:
        # The reference's pencil layout (config.py:194-195; spectral arrays (N0, N1/P0, Nh/P1)) is not built on the
        # B200 path: on one NVSwitch box a slab needs one exchange per transform where a pencil needs two.  Silently
        # running slab would hand the caller arrays of a different local shape, so refuse.
        raise NotImplementedError("spectraldns_b200: --decomposition pencil on %d GPUs is not implemented "
                                  "(slab is; on a single GPU the flag is accepted because both coincide)" % comm.Get_size())
    eng = Engine.get(params.N, params.L, params.precision, params.dealias, solver_name,
                     params.mask_nyquist, params.decomposition, conv)
    T = TensorProductSpace(comm, V, dtype=float_, slab=(params.decomposition == 'slab'),
                           engine=eng, which=0, solver=solver_name, mask_nyquist=params.mask_nyquist)
    Tp = T.get_dealiased(padding_factor=1.5 if params.dealias == '3/2-rule' else 1,
                         dealias_direct=params.dealias == '2/3-rule')
    Tp._engine = eng
    if params.dealias == 'None':
        Tp._which = 0
    return V, T, Tp, eng


def wavenumber_arrays(T, VT, float_):
    """X, K, K2, K_over_K2 exactly as get_context builds them (solvers/NS.py:36-48)."""
    dim = len(T.N)
    X = T.local_mesh(True)
    K = T.local_wavenumbers(scaled=True)
    for i in range(dim):
        X[i] = X[i].astype(float_)
        K[i] = K[i].astype(float_)
    K2 = np.zeros(T.shape(True), dtype=float_)
    for i in range(dim):
        K2 += K[i]*K[i]
    K_over_K2 = np.zeros(VT.shape(True), dtype=float_)
    for i in range(dim):
        K_over_K2[i] = K[i] / np.where(K2 == 0, 1, K2)
    return X, K, K2, K_over_K2


def pinned_like(shape, dtype):
    """Host array in page-locked memory (fast, asynchronous H2D/D2H of the solver state)."""
    import torch
    tdt = {np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128,
           np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
    t = torch.zeros(tuple(shape), dtype=tdt, pin_memory=True)
    a = t.numpy()
    return a, t


def device_state(context):
    """The DeviceState of a context (created on first use)."""
    dev = context.get('_dev', None)
    if dev is None or not dev.is_state(context['u']):
        dev = DeviceState(context['_engine'], context['u'], int(context['u'].shape[0]))
        dict.__setitem__(context, '_dev', dev)
        hf = context.get('hdf5file', None)
        if hf is not None:
            hf.before_host_read = dev.sync_to_host
    return dev


def run_rhs(dev, rhs, u_hat, source, want_p):
    """ComputeRHS on host arrays: stage in, five kernel launches, stage out."""
    params = config.params
    plan = dev.plan
    if plan.solver == 'NS' and params.convection != plan.convection:
        raise RuntimeError('params.convection changed after get_context(): the CUDA plan was built for %r' % plan.convection)
    plan.use_current_stream()
    if hasattr(plan, 'set_physics'):
        plan.set_physics(params)            # Bq2D: Richardson / Prandtl numbers (config.py:260-261)
    d_u = dev.device_input(u_hat)
    src = dev.refresh_source(source)
    d_rhs = dev.rhs_buffer()
    d_p = None
    if want_p is not None:
        d_p = dev.engine.stage('p_hat', plan.spectral_shape, plan.tcomplex)
    eta = float(params.eta) if 'eta' in params else 0.0
    plan.compute_rhs(d_rhs, d_u, float(params.nu), eta, source=src, p_hat=d_p)
    rhs[...] = d_rhs.cpu().numpy()
    if d_p is not None:
        want_p[...] = d_p.cpu().numpy()
    return rhs


def dev_of(context_rest):
    """DeviceState travelling in the **context remainder of the reference-style signatures."""
    dev = context_rest.get('_dev', None)
    if dev is None:
        raise RuntimeError('ComputeRHS must be called with **context of get_context()')
    return dev


class Convection(object):
    """Callable returned by getConvection(); carries .convection like the reference's closure
    (solvers/NS.py:200).  Signature per solver: Conv(rhs, u_hat, <spaces and work arrays...>)."""
    _sdns_builtin = True

    def __init__(self, name):
        self.convection = name

    def __call__(self, rhs, u_hat, *args, **kwargs):
        eng = None
        for a in list(args) + list(kwargs.values()):
            eng = getattr(a, '_engine', None) or getattr(getattr(a, 'T', None), '_engine', None)
            if eng is not None:
                break
        if eng is None:
            raise RuntimeError('conv(): pass the dealiased space Tp/VTp of get_context()')
        p = eng.plan
        if not hasattr(p, 'compute_conv'):
            raise NotImplementedError('conv() on its own is not part of the 2-D C ABI; ComputeRHS runs it fused')
        p.use_current_stream()
        d_u = eng.upload('conv_in', u_hat, p.complex, p.tcomplex)
        d_r = eng.stage('conv_out', d_u.shape, p.tcomplex)
        p.compute_conv(d_r, d_u)
        rhs[...] = d_r.cpu().numpy()
        return rhs
