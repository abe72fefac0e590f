"""On-device diagnostics and low-wavenumber forcing for isotropic-turbulence runs.

demo/Isotropic.py computes its per-step forcing (update(): :161-184), its energy spectrum (spectrum(): :88-118) and its
dissipation / divergence monitors (:220-259) with numpy on the host arrays of the context.  On the B200 path the
solution lives on the GPU while solve() runs; these functions compute the same quantities where the state is, through
the C ABI (sdns_energy_weighted, sdns_scale_field, sdns_set_mode, sdns_spectrum, sdns_enstrophy,
sdns_divergence_norm) -- no copy of the state in either direction.

Two ways in:
  * a callback written against this module (INTEGRATION.md shows the three-line change to demo/Isotropic.py's update);
  * the UNCHANGED demo with SDNS_LAZY_STATE=1: the context's state array then defers its device-to-host copy until host
    code really reads it, and the expressions of update() -- energy_fourier(U_hat, T), energy_fourier(U_hat*k2_mask, T),
    U_hat[:, 0, 0, 0] = 0, U_hat *= factor -- are recognised and routed here (spaces._SpaceArray).
"""
import numpy as np


def _dev(context):
    dev = context.get('_dev', None) if hasattr(context, 'get') else getattr(context, '_dev', None)
    if dev is None:
        raise RuntimeError('the context has no device state yet (call solver.device_state(context) or solve())')
    if dev.host_dirty:
        dev.upload_state()
    return dev


def _allreduce(context, x):
    comm = context.T.comm
    return comm.allreduce(x) if comm is not None and hasattr(comm, 'allreduce') else x


def real_field(dev, field, tag=None):
    """Device tensor (float32 / float64, spectral shape) of a real field given as a host array or already as a CUDA
    tensor.  Host arrays are uploaded on every call -- hashing one to validate a cache costs more than the copy, and the
    copy is one real number per mode against six for the state; keep the returned tensor (device_field) to avoid it."""
    import torch
    p = dev.plan
    if isinstance(field, torch.Tensor):
        t = field
    else:
        a = np.asarray(field)
        a = np.ascontiguousarray(a) if a.shape == tuple(p.spectral_shape) else np.array(np.broadcast_to(a, p.spectral_shape))
        t = torch.from_numpy(a).to(p.device)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32 if p.precision == 'single' else torch.float64)     # integer / boolean masks: converted on the device
    return t.contiguous()


def device_field(context, host_array):
    """Upload a real field of the spectral shape (e.g. k2_mask) once; pass the result to energy / forcing_rescale."""
    return real_field(_dev(context), host_array)


def energy(context, weight=None):
    """energy_fourier(U_hat, T), or energy_fourier(U_hat*weight, T) for a real host field `weight` (Isotropic.py:167-168)."""
    dev = _dev(context)
    w = real_field(dev, weight, 'energy_weight') if weight is not None else None
    return float(_allreduce(context, dev.plan.energy_weighted(dev.u, w)))


def forcing_rescale(context, k2_mask, target_energy, zero_mean=True):
    """The forcing of demo/Isotropic.py:161-184 without leaving the GPU: zero the mean mode, rescale the modes inside
    k2_mask so that the total energy returns to target_energy.  Returns (energy_new, alpha)."""
    dev = _dev(context)
    p = dev.plan
    if zero_mean and p.rank == 0:
        p.set_mode(dev.u, (0, 0, 0), 0.0)
    mask = real_field(dev, k2_mask, 'k2_mask')
    e_all = float(_allreduce(context, p.energy_weighted(dev.u, None)))
    e_low = float(_allreduce(context, p.energy_weighted(dev.u, mask)))
    alpha = np.sqrt((target_energy - (e_all - e_low))/e_low)
    p.scale_field(dev.u, mask, alpha, 1.0)            # U_hat *= alpha*k2_mask + (1 - k2_mask)
    dev.device_newer = True
    return float(_allreduce(context, p.energy_weighted(dev.u, None))), float(alpha)


def spectrum(context):
    """(Ek, bins) of demo/Isotropic.py:88-118, shells summed on the device."""
    dev = _dev(context)
    N = np.array(context.T.N, dtype=float)
    Nb = int(np.sqrt(sum((N/2)**2)/3))
    bins = np.array(range(0, Nb)) + 0.5
    sums, counts = dev.plan.spectrum_shells(dev.u, Nb)
    sums = np.asarray(_allreduce(context, sums))
    counts = np.asarray(_allreduce(context, counts))
    Ek = np.zeros(Nb)
    for i in range(Nb-1):
        if counts[i]:
            Ek[i] = (bins[i+1]**3 - bins[i]**3)*(4./3.*np.pi)*sums[i]/counts[i]
    return Ek, bins


def dissipation(context):
    """energy_fourier(1j*K x U_hat, T) (Isotropic.py:243-244); multiply by nu for eps."""
    dev = _dev(context)
    return float(_allreduce(context, dev.plan.enstrophy(dev.u)))


def divergence_norm(context):
    """L2_norm(get_divergence(**context)) (Isotropic.py:245-247) through Parseval's identity."""
    dev = _dev(context)
    return float(_allreduce(context, dev.plan.divergence_norm(dev.u)))
