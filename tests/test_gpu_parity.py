"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against the
CPU oracle and the golden fixtures generated from the unmodified reference.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-11 in double, <= 1e-4 in single."""
import glob
import os
import numpy as np
import pytest
from conftest import golden, rel_l2, GOLDEN
import sdns_oracle as so

pytestmark = pytest.mark.gpu

TOL = {'double': 1e-11, 'single': 1e-4}
ALL = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, '*.npz'))
             if not os.path.basename(f).startswith('integ_'))


def make_plan(N, L=(2*np.pi,)*3, precision='double', dealias='2/3-rule', solver='NS', mask_nyquist=True,
              convection=None):
    from spectraldns_b200.plan import Plan
    return Plan(N, L, precision, dealias, solver, convection=convection, mask_nyquist=mask_nyquist)


def test_library_loaded_and_native():
    from spectraldns_b200 import _lib
    L = _lib.lib()
    assert L.sdns_abi_version() == 1
    assert L.sdns_size_supported(256, 1) == 0
    assert L.sdns_size_supported(100, 1) != 0


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('N', [(16, 16, 16), (32, 32, 32), (64, 32, 16), (16, 32, 64), (128, 128, 128),
                               (24, 48, 96), (192, 16, 32)])
def test_plain_transforms(N, precision):
    """T.forward / T.backward (solvers/NS.py:93,103): rfftn/prod(N) and its inverse."""
    o = so.Oracle(N, precision=precision, dealias='None')
    p = make_plan(N, precision=precision, dealias='None')
    rng = np.random.RandomState(1)
    u = rng.standard_normal((3,)+tuple(N)).astype(o.float)
    u_hat = p.to_host(p.forward(p.to_device(u)))
    ref = o.forward(u.astype(np.float64)) if precision == 'double' else o.forward(u)
    assert rel_l2(u_hat, ref) < TOL[precision]
    back = p.to_host(p.backward(p.to_device(ref.astype(o.complex))))
    assert rel_l2(back, u) < TOL[precision]
    # odd component counts and a single scalar field
    for nc in (1, 2, 5):
        v = rng.standard_normal((nc,)+tuple(N)).astype(o.float)
        vh = p.to_host(p.forward(p.to_device(v)))
        assert rel_l2(vh, o.forward(v)) < TOL[precision]
        vb = p.to_host(p.backward(p.to_device(vh)))
        assert rel_l2(vb, v) < 10*TOL[precision]


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('dealias', ['2/3-rule', '3/2-rule'])
@pytest.mark.parametrize('N', [(16, 16, 16), (32, 16, 64), (64, 64, 64)])
def test_dealiased_space_transforms(N, dealias, precision):
    """Tp.backward / Tp.forward (solvers/NS.py:29-32): truncated input or 3/2 padding."""
    o = so.Oracle(N, precision=precision, dealias=dealias)
    p = make_plan(N, precision=precision, dealias=dealias)
    rng = np.random.RandomState(2)
    u_hat = o.forward(rng.standard_normal((3,)+tuple(N)).astype(o.float))
    ref = o._bwd_p(u_hat)
    got = p.to_host(p.backward(p.to_device(u_hat), padded=True))
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < TOL[precision]
    w = rng.standard_normal(ref.shape).astype(o.float)
    got_f = p.to_host(p.forward(p.to_device(w), padded=True))
    assert rel_l2(got_f, o._fwd_p(w)) < TOL[precision]


def _state0(o, g):
    if 'u0_hat' in g:
        return g['u0_hat']
    if str(g['solver']) == 'MHD':
        return o.forward(so.taylor_green_mhd(o))
    u = o.forward(so.taylor_green(o))
    if str(g['solver']) == 'VV':
        u = o.cross2(o.K, u)
    return u


@pytest.mark.parametrize('name', ALL)
def test_golden_rhs_and_rk4(name):
    """ComputeRHS and N RK4 steps against fixtures produced by the unmodified reference
    (oracle/make_golden.py): velocity (state) field, relative L2."""
    g = golden(name)
    prec, solver = str(g['precision']), str(g['solver'])
    mask = 'nodealias' not in name
    o = so.Oracle(g['N'], g['L'], prec, str(g['dealias']), mask_nyquist=mask)
    p = make_plan(g['N'], g['L'], prec, str(g['dealias']), solver, mask_nyquist=mask)
    u0 = _state0(o, g).astype(o.complex)
    eta = float(g['eta']) if 'eta' in g else 0.0
    conv = 'Divergence' if solver == 'MHD' else 'Vortex'
    if 'rhs_'+conv in g.files:
        d_u = p.to_device(u0)
        rhs = p.compute_rhs(p.empty_spectral(), d_u, float(g['nu']), eta)
        assert rel_l2(p.to_host(rhs), g['rhs_'+conv]) < TOL[prec], 'rhs'
        assert rel_l2(p.to_host(d_u), u0) == 0.0          # input untouched
    d_u = p.to_device(u0)
    u1, u2 = p.empty_spectral(), p.empty_spectral()
    for _ in range(int(g['nsteps'])):
        p.rk4_step(d_u, u1, u2, float(g['dt']), float(g['nu']), eta)
    err = rel_l2(p.to_host(d_u), g['u_hat'])
    assert err < TOL[prec], err


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('dealias', ['2/3-rule', '3/2-rule'])
@pytest.mark.parametrize('conv', ['Standard', 'Divergence', 'Skewed'])
def test_ns_convection_forms(conv, dealias, precision):
    """NS.getConvection 'Standard' / 'Divergence' / 'Skewed' (solvers/NS.py:138-189): ComputeRHS against the
    reference fixture (2/3-rule, double) and the oracle, plus two RK4 steps."""
    g = golden('iso_ns_16_double')
    N = tuple(int(n) for n in g['N'])
    o = so.Oracle(N, g['L'], precision, dealias)
    p = make_plan(N, g['L'], precision, dealias, 'NS', convection=conv)
    u0 = g['u0_hat'].astype(o.complex)
    nu = float(g['nu'])
    rhs = p.to_host(p.compute_rhs(p.empty_spectral(), p.to_device(u0), nu))
    assert rel_l2(rhs, o.ns_rhs(u0, nu, conv)) < TOL[precision]
    if dealias == '2/3-rule' and precision == 'double':
        assert rel_l2(rhs, g['rhs_' + conv]) < 1e-11            # the unmodified reference
    cv = p.to_host(p.compute_conv(p.empty_spectral(), p.to_device(u0)))
    assert rel_l2(cv, o.ns_conv(u0, conv)) < TOL[precision]
    d_u, u1, u2 = p.to_device(u0), p.empty_spectral(), p.empty_spectral()
    for _ in range(2):
        p.rk4_step(d_u, u1, u2, 0.005, nu)
    ref = o.solve(u0, 'NS', 2, 0.005, nu, convection=conv)
    assert rel_l2(p.to_host(d_u), ref) < TOL[precision]


def test_pressure_and_source():
    g = golden('iso_ns_16_double')
    o = so.Oracle(g['N'], g['L'], 'double', '2/3-rule')
    p = make_plan(g['N'], g['L'], 'double', '2/3-rule', 'NS')
    rng = np.random.RandomState(5)
    src = (rng.standard_normal(g['u0_hat'].shape) + 1j*rng.standard_normal(g['u0_hat'].shape))*1e-2
    ref, P = o.ns_rhs(g['u0_hat'], float(g['nu']), source=src, return_p=True)
    p_hat = p.empty_spectral(0)
    rhs = p.compute_rhs(p.empty_spectral(), p.to_device(g['u0_hat']), float(g['nu']),
                        source=p.to_device(src), p_hat=p_hat)
    assert rel_l2(p.to_host(rhs), ref) < 1e-11
    assert rel_l2(p.to_host(p_hat), P) < 1e-11


@pytest.mark.parametrize('precision', ['double', 'single'])
def test_tg_known_answers_on_gpu(precision):
    """tests/TG.py:116-126 of the reference at 32^3 (BASELINE.json configs[0]) through the GPU path:
    k and w after 10 RK4 steps."""
    N = (32, 32, 32)
    o = so.Oracle(N, precision=precision)
    p = make_plan(N, precision=precision)
    U = p.to_device(so.taylor_green(o))
    u = p.forward(U)
    u1, u2 = p.empty_spectral(), p.empty_spectral()
    for _ in range(10):
        p.rk4_step(u, u1, u2, 0.01, 0.000625)
    Uf = p.to_host(p.backward(u)).astype(np.float64)
    curl_hat = p.cross2(p.empty_spectral(), u)
    W = p.to_host(p.backward(curl_hat)).astype(np.float64)
    k = np.sum(Uf*Uf)/np.prod(N)/2
    w = np.sum(W*W)/np.prod(N)/2
    nt = 7 if precision == 'double' else 5
    assert round(float(w) - 0.375249930801, nt) == 0
    assert round(float(k) - 0.124953117517, nt) == 0
    # Parseval (tests/TG.py:101-109)
    assert abs(p.energy(u)/2 - k) < (1e-12 if precision == 'double' else 1e-6)


def test_integrators_euler_ab2():
    """maths/integrators.py:161-175 through sdns_euler_step / sdns_ab2_step."""
    g = golden('iso_ns_16_double')
    o = so.Oracle(g['N'], g['L'], 'double', '2/3-rule')
    p = make_plan(g['N'], g['L'], 'double', '2/3-rule', 'NS')
    nu, dt = float(g['nu']), float(g['dt'])
    fn = lambda u: o.ns_rhs(u, nu)
    ref = g['u0_hat'].copy()
    d_u, rhs = p.to_device(g['u0_hat']), p.empty_spectral()
    for _ in range(3):
        ref = o.forward_euler_step(ref, fn, dt)
        p.euler_step(d_u, rhs, dt, nu)
    assert rel_l2(p.to_host(d_u), ref) < 1e-11
    ref, r1 = g['u0_hat'].copy(), np.zeros_like(g['u0_hat'])
    d_u, u1 = p.to_device(g['u0_hat']), p.empty_spectral()
    for ts in range(3):
        ref, r1 = o.ab2_step(ref, r1, fn, dt, ts)
        p.ab2_step(d_u, u1, rhs, dt, ts, nu)
    assert rel_l2(p.to_host(d_u), ref) < 1e-11


@pytest.mark.parametrize('cfg', [((256, 256, 256), 'double', '2/3-rule'),
                                 ((128, 128, 128), 'single', '3/2-rule'),
                                 ((512, 64, 128), 'double', '2/3-rule')])
def test_full_size_properties(cfg):
    """Size-independent properties at sizes the oracle does not finish in seconds:
    forward(backward(x)) == x, Parseval, RK4 keeps the field solenoidal, energy decays."""
    N, prec, dealias = cfg
    import torch
    p = make_plan(N, precision=prec, dealias=dealias)
    tol = TOL[prec]
    g = torch.Generator(device='cuda').manual_seed(0)
    u = torch.randn((3,)+tuple(N), dtype=p.tfloat, device='cuda', generator=g)
    u_hat = p.forward(u)
    back = p.backward(u_hat)
    assert float((back-u).norm()/u.norm()) < tol
    e_phys = float((u.double()**2).sum())/np.prod(N)
    assert abs(p.energy(u_hat) - e_phys)/e_phys < 10*tol
    # TG step: analytic field, energy must decay and stay divergence free
    o = so.Oracle((16, 16, 16))        # only for the mesh formula
    X = [torch.arange(n, dtype=torch.float64, device='cuda')*2*np.pi/n for n in N]
    U = torch.zeros((3,)+tuple(N), dtype=p.tfloat, device='cuda')
    U[0] = (torch.sin(X[0])[:, None, None]*torch.cos(X[1])[None, :, None]*torch.cos(X[2])[None, None, :]).to(p.tfloat)
    U[1] = (-torch.cos(X[0])[:, None, None]*torch.sin(X[1])[None, :, None]*torch.cos(X[2])[None, None, :]).to(p.tfloat)
    uh = p.forward(U)
    e0 = p.energy(uh)
    assert abs(e0/2 - 0.125) < 1e-6
    u1, u2 = p.empty_spectral(), p.empty_spectral()
    for _ in range(2):
        p.rk4_step(uh, u1, u2, 0.01, 0.000625)
    e1 = p.energy(uh)
    assert 0 < e1 < e0
    # analytic TG decay rate at t=0: dk/dt = -2 nu w = -2*nu*0.375 -> k(0.02) ~ 0.125 - 9.4e-6
    assert abs(e1/2 - (0.125 - 2*0.000625*0.375*0.02)) < 1e-6
    k = [torch.fft.fftfreq(N[0], 1./N[0], device='cuda'), torch.fft.fftfreq(N[1], 1./N[1], device='cuda'),
         torch.fft.rfftfreq(N[2], 1./N[2], device='cuda')]
    div = (k[0][:, None, None]*uh[0] + k[1][None, :, None]*uh[1] + k[2][None, None, :]*uh[2])
    assert float(div.abs().max()) < (1e-12 if prec == 'double' else 1e-5)


def test_field_parity_at_baseline_size():
    """BASELINE configs[1]'s grid (256^3 fp64, 2/3-rule): right-hand side and one RK4 step of a seeded broadband field
    against the oracle, field level (the oracle needs about half a minute of host time at this size)."""
    N = (256, 256, 256)
    o = so.Oracle(N)
    p = make_plan(N)
    f0 = so.isotropic_field(o, seed=3)
    d_u = p.to_device(f0)
    rhs = p.to_host(p.compute_rhs(p.empty_spectral(), d_u, 0.005))
    assert rel_l2(rhs, o.ns_rhs(f0, 0.005)) < 1e-11
    u1, u2 = p.empty_spectral(), p.empty_spectral()
    p.rk4_step(d_u, u1, u2, 0.002, 0.005)
    assert rel_l2(p.to_host(d_u), o.solve(f0, 'NS', 1, 0.002, 0.005)) < 1e-11


@pytest.mark.parametrize('cfg', [((16, 16, 512), 'double', '2/3-rule'), ((16, 16, 512), 'double', '3/2-rule'),
                                 ((16, 16, 1024), 'double', '2/3-rule'), ((16, 16, 2048), 'single', '2/3-rule'),
                                 ((16, 16, 1024), 'single', '3/2-rule'), ((512, 16, 16), 'double', '2/3-rule'),
                                 ((16, 512, 16), 'double', '3/2-rule'), ((16, 16, 512), 'double', 'None'),
                                 # 1024 / 2048 (and their 3/2 paddings 1536 / 3072) on the strided axes: what the
                                 # 1024^3 and 2048^3 configurations of BASELINE.json run on axes 0 and 1
                                 ((1024, 16, 16), 'double', '2/3-rule'), ((16, 1024, 16), 'double', '2/3-rule'),
                                 ((2048, 16, 16), 'single', '2/3-rule'), ((16, 2048, 16), 'single', '2/3-rule'),
                                 ((1024, 16, 16), 'single', '3/2-rule'), ((16, 1024, 16), 'double', '3/2-rule'),
                                 ((2048, 8, 16), 'single', '3/2-rule'), ((2048, 16, 16), 'double', '2/3-rule')])
def test_long_axis_rhs(cfg):
    """Transform lengths 512..2048 on one axis (the multi-warp-per-line fused z kernel and the long strided
    passes) at a size the oracle finishes in seconds: NS ComputeRHS + one RK4 step on a random field."""
    N, prec, dealias = cfg
    o = so.Oracle(N, precision=prec, dealias=dealias)
    p = make_plan(N, precision=prec, dealias=dealias)
    rng = np.random.RandomState(7)
    u0 = o.forward(rng.standard_normal((3,)+tuple(N)).astype(o.float)*0.3).astype(o.complex)
    nu = 0.01
    rhs = p.to_host(p.compute_rhs(p.empty_spectral(), p.to_device(u0), nu))
    assert rel_l2(rhs, o.ns_rhs(u0, nu)) < TOL[prec]
    d_u, u1, u2 = p.to_device(u0), p.empty_spectral(), p.empty_spectral()
    p.rk4_step(d_u, u1, u2, 0.001, nu)
    assert rel_l2(p.to_host(d_u), o.solve(u0, 'NS', 1, 0.001, nu)) < TOL[prec]


@pytest.mark.parametrize('cfg', [((60, 60, 60), 'double', '3/2-rule', 'NS'), ((60, 60, 60), 'single', '2/3-rule', 'NS'),
                                 ((60, 16, 32), 'double', '2/3-rule', 'VV'), ((16, 90, 16), 'double', 'None', 'NS'),
                                 ((32, 16, 60), 'single', '3/2-rule', 'VV')])
def test_lengths_60_and_90(cfg):
    """demo/Isotropic.py's default grid (60^3, 3/2-rule -> 90^3 padded): NS / VV ComputeRHS + one RK4 step + transforms."""
    N, prec, dealias, solver = cfg
    o = so.Oracle(N, precision=prec, dealias=dealias)
    p = make_plan(N, precision=prec, dealias=dealias, solver=solver)
    rng = np.random.RandomState(7)
    u0 = o.forward(rng.standard_normal((3,)+tuple(N)).astype(o.float)*0.3).astype(o.complex)
    if solver == 'VV':
        u0 = o.cross2(o.K, u0).astype(o.complex)
    nu = 0.01
    ref = o.ns_rhs(u0, nu) if solver == 'NS' else o.vv_rhs(u0, nu)
    assert rel_l2(p.to_host(p.compute_rhs(p.empty_spectral(), p.to_device(u0), nu)), ref) < TOL[prec]
    d_u, u1, u2 = p.to_device(u0), p.empty_spectral(), p.empty_spectral()
    p.rk4_step(d_u, u1, u2, 0.001, nu)
    assert rel_l2(p.to_host(d_u), o.solve(u0, solver, 1, 0.001, nu)) < TOL[prec]
    u = rng.standard_normal((3,)+tuple(N)).astype(o.float)
    assert rel_l2(p.to_host(p.forward(p.to_device(u))), o.forward(u)) < TOL[prec]



def _ref_cython(precision):
    """The reference's own compiled kernels (oracle/_ref, built in the container from the Cython templates where
    they lie); None on a box where they are missing."""
    import importlib
    import sys
    import build_ref_cython as brc
    if not brc.available() and not brc.build():
        return None
    if brc.OUT not in sys.path:
        sys.path.insert(0, brc.OUT)
    return [importlib.import_module('cython_%s_%s' % (precision, m)) for m in ('maths', 'solvers')]


@pytest.mark.parametrize('precision', ['double', 'single'])
def test_standalone_operators_against_reference_cython(precision):
    """The fine-grained plug-in surface (optimization/__init__.py:12-55): sdns_cross1, sdns_cross2 (wavenumber form,
    with and without 1/k^2), sdns_cross2_dense, sdns_project and sdns_add_pressure_diffusion as stand-alone calls,
    against the reference's compiled Cython (cython_maths.in:8-86, cython_solvers.in:40-80) when oracle/_ref is
    there, and always against the oracle."""
    N = (16, 12, 24)
    Lbox = (2*np.pi, 4*np.pi, 6*np.pi)
    o = so.Oracle(N, L=Lbox, precision=precision)
    p = make_plan(N, L=Lbox, precision=precision)
    tol = 1e-14 if precision == 'double' else 1e-6
    rng = np.random.RandomState(5)
    ref = _ref_cython(precision)

    def cplx(shape):
        return (rng.standard_normal(shape) + 1j*rng.standard_normal(shape)).astype(o.complex)
    a = rng.standard_normal((3,)+N).astype(o.float)
    b = rng.standard_normal((3,)+N).astype(o.float)
    c = p.to_host(p.cross1(p.empty_physical(), p.to_device(a), p.to_device(b)))
    assert rel_l2(c, o.cross1(a, b)) < tol
    bh = cplx((3,)+o.sshape)
    c2 = p.to_host(p.cross2(p.empty_spectral(), p.to_device(bh)))
    assert rel_l2(c2, o.cross2(o.K, bh)) < tol
    c2k = p.to_host(p.cross2(p.empty_spectral(), p.to_device(bh), over_k2=True))
    assert rel_l2(c2k, o.cross2(o.K_over_K2, bh)) < tol
    ad = rng.standard_normal((3,)+o.sshape).astype(o.float)
    cd = p.to_host(p.cross2_dense(p.empty_spectral(), p.to_device(ad), p.to_device(bh)))
    assert rel_l2(cd, o.cross2(ad, bh)) < tol
    pr = p.to_host(p.project(p.to_device(bh)))
    pref = bh - np.sum(o.K_over_K2*bh, 0)*np.array([np.broadcast_to(k, o.sshape) for k in o.K])
    assert rel_l2(pr, pref) < tol
    du, uh = cplx((3,)+o.sshape), cplx((3,)+o.sshape)
    nu = 0.0123
    d_du, d_p = p.to_device(du), p.empty_spectral(0)
    p.add_pressure_diffusion(d_du, p.to_device(uh), nu, d_p)
    du_or, p_or = o.add_pressure_diffusion(du.copy(), uh, o.float(nu))
    assert rel_l2(p.to_host(d_du), du_or) < tol and rel_l2(p.to_host(d_p), p_or) < tol
    if ref is not None:
        maths, solvers = ref
        assert rel_l2(c, maths.cross1(np.zeros_like(a), a, b)) < tol
        assert rel_l2(c2, maths.cross2(np.zeros_like(bh), [np.ascontiguousarray(k) for k in o.K], bh)) < tol
        assert rel_l2(cd, maths.cross2(np.zeros_like(bh), ad, bh)) < tol
        p_ref = np.zeros(o.sshape, dtype=o.complex)
        du_ref = solvers.add_pressure_diffusion_NS(du.copy(), uh, o.float(nu), o.K2, o.K, p_ref, o.K_over_K2)
        assert rel_l2(p.to_host(d_du), du_ref) < tol and rel_l2(p.to_host(d_p), p_ref) < tol


@pytest.mark.parametrize('precision', ['double', 'single'])
def test_device_diagnostics_match_the_demo(precision):
    """sdns_spectrum / sdns_enstrophy / sdns_divergence_norm / sdns_energy_weighted / sdns_scale_field / sdns_set_mode
    against the numpy of demo/Isotropic.py (:88-118 spectrum, :161-184 forcing, :243-247 monitors) restated in the
    oracle, on a seeded isotropic field."""
    import torch
    N = (32, 48, 64)
    o = so.Oracle(N, precision=precision)
    p = make_plan(N, precision=precision)
    tol = 1e-12 if precision == 'double' else 2e-6
    u0 = so.isotropic_field(o, seed=3)
    d_u = p.to_device(u0)
    Ek_ref, bins = o.spectrum(u0)
    sums, cnts = p.spectrum_shells(d_u, len(bins))
    Ek = np.zeros(len(bins))
    for i in range(len(bins)-1):
        if cnts[i]:
            Ek[i] = (bins[i+1]**3 - bins[i]**3)*(4./3.*np.pi)*sums[i]/cnts[i]
    assert np.allclose(Ek, Ek_ref, rtol=tol, atol=0) and Ek_ref.max() > 0
    assert abs(p.enstrophy(d_u) - o.enstrophy(u0)) < tol*o.enstrophy(u0)
    ud = (u0 + o.forward(np.random.RandomState(1).standard_normal((3,)+N))*o.mask*1e-3).astype(o.complex)   # not solenoidal
    assert abs(p.divergence_norm(p.to_device(ud)) - o.divergence_norm(ud)) < 20*tol*o.divergence_norm(ud)
    # the forcing step, expression by expression
    g = u0.copy()
    g[:, 0, 0, 0] = 0.3 - 0.1j
    d_g = p.to_device(g)
    k2_mask = np.where(o.K2 <= 3**2, 1, 0)
    target = 1.05*o.energy_fourier(u0)
    ref, e_ref, e_low, alpha = o.forcing_rescale(g.copy(), 3, target)
    p.set_mode(d_g, (0, 0, 0), 0.0)
    d_m = torch.from_numpy(k2_mask.astype(np.float64)).cuda()
    assert abs(p.energy_weighted(d_g, d_m) - e_low) < tol*e_low
    p.scale_field(d_g, d_m, alpha, 1.0)
    assert rel_l2(p.to_host(d_g), ref) < (1e-15 if precision == 'double' else 2e-7)
    assert abs(p.energy_weighted(d_g) - e_ref) < tol*e_ref and abs(p.energy(d_g) - e_ref) < tol*e_ref
    # float32 factor array, plain product
    f32 = torch.from_numpy((alpha*k2_mask + (1-k2_mask)).astype(np.float32)).cuda()
    d_h = p.to_device(g)
    p.set_mode(d_h, (0, 0, 0), 0.0)
    p.scale_field(d_h, f32)
    assert rel_l2(p.to_host(d_h), ref) < 2e-7
