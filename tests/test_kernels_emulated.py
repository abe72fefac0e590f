"""Kernel logic without a GPU: the library's own CUDA sources (plan, C ABI, every pass kernel) compiled by g++
against tests/host/host_shim.h -- one OS thread per CUDA thread, barriers for __syncthreads / __syncwarp, mailboxes
for warp shuffles, NaN-filled shared memory and workspace -- and run on tiny grids through the same C ABI calls as
the GPU parity tests, against the CPU oracle.  This checks indexing, axis maps, truncation / padding, exchange maps,
epilogues and stage updates of the PRODUCT kernels; speed and anything that depends on real hardware (memory model,
occupancy) stay with the -m gpu tests.  The emulated library is test infrastructure: the product never loads it."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'tests', 'host'), os.path.join(ROOT, 'oracle')]
import sdns_oracle as so          # noqa: E402
from conftest import rel_l2      # noqa: E402

TOL = {'double': 1e-11, 'single': 1e-4}
# The emulated multi-rank matrix takes a quarter of an hour on 8 cores.  The default run keeps one representative of every
# mechanism (each exchange mode, each rank count, both axis-1 layouts, the skewed-rank race detector); SDNS_EMU_FULL=1 runs
# all of it (what every change to the exchange was developed against).
FULL = os.environ.get('SDNS_EMU_FULL') == '1'
full_only = pytest.mark.skipif(not FULL, reason='part of the full emulated matrix: SDNS_EMU_FULL=1')


@pytest.fixture(scope='module')
def emu():
    os.environ.setdefault('SDNS_EMU_JITTER', '100')      # random start delay per launch (us): pulls the ranks apart
    import build_emu
    import emu_plan
    return emu_plan.load(build_emu.build()), emu_plan


def _state(o, solver, seed=3):
    f0 = so.isotropic_field(o, seed=seed, ncomp=6 if solver == 'MHD' else 3)
    if solver == 'VV':
        f0 = o.cross2(o.K, f0)
    return f0.astype(o.complex)


@pytest.mark.parametrize('precision', ['double', 'single'])
@pytest.mark.parametrize('N', [(16, 16, 16), (8, 12, 24), (32, 16, 8), (24, 48, 16)])
def test_emulated_plain_transforms(emu, N, precision):
    L, ep = emu
    o = so.Oracle(N, precision=precision, dealias='None')
    p = ep.EmuPlan(L, N, precision=precision, dealias='None')
    rng = np.random.RandomState(1)
    u = rng.standard_normal((3,)+tuple(N)).astype(o.float)
    assert rel_l2(p.forward(u), o.forward(u)) < TOL[precision]
    assert rel_l2(p.backward(o.forward(u).astype(o.complex)), u) < TOL[precision]
    p.close()


@pytest.mark.parametrize('solver,N,dealias,precision', [
    ('NS', (16, 16, 16), '2/3-rule', 'double'), ('NS', (16, 16, 16), '3/2-rule', 'single'), ('NS', (32, 16, 8), '3/2-rule', 'double'),
    ('NS', (32, 16, 8), 'None', 'single'), ('NS', (8, 24, 48), '2/3-rule', 'single'),
    ('VV', (16, 32, 16), '2/3-rule', 'double'), ('VV', (16, 32, 16), '3/2-rule', 'single'), ('VV', (16, 16, 16), 'None', 'double'),
    ('MHD', (16, 16, 32), '2/3-rule', 'double'), ('MHD', (16, 16, 32), '3/2-rule', 'double'), ('MHD', (16, 16, 16), '2/3-rule', 'single')])
def test_emulated_rhs_and_rk4(emu, solver, N, dealias, precision):
    L, ep = emu
    o = so.Oracle(N, precision=precision, dealias=dealias)
    p = ep.EmuPlan(L, N, precision=precision, dealias=dealias, solver=solver)
    f0 = _state(o, solver)
    nu, eta, dt = 0.005, 0.01, 0.002
    ref = {'NS': lambda: o.ns_rhs(f0, nu), 'VV': lambda: o.vv_rhs(f0, nu), 'MHD': lambda: o.mhd_rhs(f0, nu, eta)}[solver]()
    assert rel_l2(p.compute_rhs(f0, nu, eta), ref) < TOL[precision]
    got = p.rk4(f0, 2, dt, nu, eta)
    assert rel_l2(got, o.solve(f0, solver, 2, dt, nu, eta=eta)) < TOL[precision]
    p.close()


@pytest.mark.parametrize('conv', ['Standard', 'Divergence', 'Skewed'])
def test_emulated_ns_convection_forms(emu, conv):
    L, ep = emu
    N = (16, 16, 16)
    o = so.Oracle(N)
    p = ep.EmuPlan(L, N, convection=conv)
    f0 = _state(o, 'NS')
    assert rel_l2(p.compute_rhs(f0, 0.005), o.ns_rhs(f0, 0.005, conv)) < 1e-11
    p.close()


# ---------------------------------------------------------------------------------------------
# multi-GPU schedule: the ranks are threads of this process, "peer memory" is each other's host buffer, the
# device-side flag barrier spins on real flags, the copy-engine copies are memcpys at the point of submission
# ---------------------------------------------------------------------------------------------
def _multi_case(emu, world, exchange, case, chunks='4', skew=None, k1_layout='blocks'):
    L, ep = emu
    N, prec, dealias, solver = case[:4]
    kcut = case[4] if len(case) > 4 else None
    conv = case[5] if len(case) > 5 else None
    os.environ['SDNS_EXCHANGE'], os.environ['SDNS_CHUNKS'] = exchange, chunks
    tol = TOL[prec]
    o = so.Oracle(N, precision=prec, dealias=dealias, kcut=kcut)
    f0 = _state(o, solver)
    nu, eta, dt = 0.005, 0.01, 0.002
    r_ref = {'NS': lambda: o.ns_rhs(f0, nu, conv or 'Vortex'), 'VV': lambda: o.vv_rhs(f0, nu),
             'MHD': lambda: o.mhd_rhs(f0, nu, eta)}[solver]()
    s_ref = o.solve(f0, solver, 1, dt, nu, eta=eta, **({'convection': conv} if conv else {}))
    rng = np.random.RandomState(11)
    u = rng.standard_normal((3,)+tuple(N)).astype(o.float)
    uh_ref = o.forward(u)

    def rank_fn(rank, sync):
        L.sdns_emu_set_skew(int(skew(rank)) if skew else 0)       # this rank's delay before every launch (microseconds)
        p = ep.EmuPlan(L, N, precision=prec, dealias=dealias, solver=solver, convection=conv, kcut=kcut, rank=rank, nranks=world,
                       k1_layout=k1_layout)
        p.open_peers(sync(p.handle()))
        N1l = N[1]//world
        k1s, x0s = p.k1_slice, p.x0_slice
        if N[1] % world == 0:
            assert k1s == (slice(rank, N[1], world) if k1_layout == 'cyclic' else slice(rank*N1l, (rank+1)*N1l))
        # the sequence visits every transition between operations that store into the peers
        e = [rel_l2(p.forward(u[:, x0s]), uh_ref[:, :, k1s]),
             rel_l2(p.backward(uh_ref[:, :, k1s].astype(o.complex)), u[:, x0s]),
             rel_l2(p.backward(uh_ref[:, :, k1s].astype(o.complex)), u[:, x0s]),
             rel_l2(p.compute_rhs(f0[:, :, k1s], nu, eta), r_ref[:, :, k1s]),
             rel_l2(p.forward(u[:, x0s]), uh_ref[:, :, k1s]),
             rel_l2(p.forward(u[:, x0s]), uh_ref[:, :, k1s]),
             rel_l2(p.rk4(f0[:, :, k1s], 1, dt, nu, eta), s_ref[:, :, k1s]),
             rel_l2(p.backward(uh_ref[:, :, k1s].astype(o.complex)), u[:, x0s])]
        assert not p.timed_out()
        sync(None)
        p.close()
        return e
    try:
        results = ep.run_ranks(world, rank_fn)
    finally:
        os.environ.pop('SDNS_EXCHANGE', None)
        os.environ.pop('SDNS_CHUNKS', None)
    for rank, e in enumerate(results):
        assert all(x < tol for x in e), (world, exchange, case, rank, e)


MULTI = [((16, 16, 16), 'double', '2/3-rule', 'NS'), ((16, 16, 16), 'double', '3/2-rule', 'NS'),
         ((32, 16, 8), 'single', '2/3-rule', 'VV'), ((16, 32, 16), 'double', '2/3-rule', 'MHD'),
         ((16, 16, 16), 'double', 'None', 'NS'), ((16, 32, 16), 'double', '2/3-rule', 'NS', (-1, 3, -1)),
         ((16, 16, 16), 'double', '2/3-rule', 'NS', None, 'Skewed'), ((16, 16, 16), 'double', '3/2-rule', 'NS', None, 'Standard')]


# (world, cases): every case runs somewhere, 8 ranks get the ones with ranks that own no kept mode
PICK = {2: (2, 6), 4: (3, 7), 8: (1, 5)}


@pytest.mark.parametrize('exchange', ['tma', 'ce', 'store'])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_emulated_multi_gpu_schedule(emu, world, exchange):
    """'tma' (the default): send slots moved by the transfer role inside the following pass kernels (csrc/xfer.cuh;
    the bulk-async copies are memcpys here, the ring / piece bookkeeping runs unchanged)."""
    if not FULL and world > 2 and exchange != 'tma':
        pytest.skip('default run: the copy-engine and peer-store modes on 2 ranks, the transfer role on 2, 4 and 8')
    for i in PICK[world][:1 if (world == 8 and (exchange == 'store' or not FULL)) else None]:
        _multi_case(emu, world, exchange, MULTI[i], chunks='6' if exchange == 'tma' else '4')


@pytest.mark.parametrize('world', [2, 4, 8])
def test_emulated_multi_gpu_cyclic_k1(emu, world):
    """Plan(k1_layout='cyclic'): rank r owns the axis-1 modes [r::P], so that the 2/3 rule leaves every rank the same number
    of kept modes.  Same global results; every exchange mode, the three dealias modes, the convection forms, MHD."""
    cases = {2: (('tma', 0), ('ce', 7), ('store', 3)), 4: (('tma', 1), ('ce', 5), ('store', 6), ('tma', 3)),
             8: (('tma', 5), ('ce', 0), ('tma', 2))}[world]
    for exchange, i in (cases if FULL else cases[:2 if world == 2 else 1]):
        _multi_case(emu, world, exchange, MULTI[i], chunks='6' if exchange == 'tma' else '4', k1_layout='cyclic')
    # the axis-1 passes find their rows in closed form when the ranks divide the threads of a line (64 / 8 = 8 threads for
    # 8 ranks, 48 / 12 = 4 for 4 ranks and the padded length), through the tables otherwise (the 8-rank cases above)
    if world == 8:
        _multi_case(emu, 8, 'tma', ((16, 64, 8), 'double', '2/3-rule', 'NS'), chunks='3', k1_layout='cyclic')
    if world == 4:
        _multi_case(emu, 4, 'tma', ((16, 32, 8), 'double', '3/2-rule', 'NS'), chunks='3', k1_layout='cyclic')


@pytest.mark.parametrize('exchange', ['tma', 'ce', 'store'])
def test_emulated_multi_gpu_uneven_slabs(emu, exchange):
    """Grid extents the rank count does not divide: N // P entries per rank and one more on the first N % P ranks
    (mpi4py-fft's slabs), for the spectral axis 1, the physical axis 0 and its 3/2-padded length."""
    if not FULL and exchange == 'ce':
        pytest.skip('default run: the transfer role and the fused peer stores')
    chunks = '3' if exchange == 'tma' else '2'
    # 3 ranks: 16 = 6 + 5 + 5 on both axes, kept axis-1 modes 0..5 | 11..15;  padded 24 = 8 + 8 + 8 with N1 uneven
    _multi_case(emu, 3, exchange, ((16, 16, 8), 'double', '2/3-rule', 'NS'), chunks=chunks)
    _multi_case(emu, 3, exchange, ((16, 16, 8), 'double', '3/2-rule', 'NS'), chunks=chunks)
    # 8 ranks: 12 = 2,2,2,2,1,1,1,1 on both axes; and the padded axis 0 alone uneven (8 -> 12 planes)
    _multi_case(emu, 8, exchange, ((12, 12, 8), 'double', '2/3-rule', 'NS'), chunks=chunks)
    if FULL or exchange == 'tma':
        _multi_case(emu, 8, exchange, ((8, 16, 8), 'double', '3/2-rule', 'NS'), chunks=chunks)
    if FULL:
        _multi_case(emu, 5, exchange, ((16, 24, 8), 'single', '2/3-rule', 'MHD'), chunks=chunks)
        _multi_case(emu, 3, exchange, ((16, 16, 8), 'double', 'None', 'NS', None, 'Skewed'), chunks=chunks)


def test_emulated_multi_gpu_transfer_role_variants(emu):
    """plain load / store transfer role; a budget so small that most of the exchange ends up in transfer-only launches;
    one so large that the first pass after a chunk carries all of it."""
    _multi_case(emu, 4, 'ldst', MULTI[3], chunks='3')
    for ratio in (('0.01', '10') if FULL else ()):
        os.environ['SDNS_XRATIO'] = ratio
        try:
            _multi_case(emu, 4, 'tma', MULTI[1], chunks='5')
        finally:
            os.environ.pop('SDNS_XRATIO', None)


@pytest.mark.parametrize('exchange', ['tma', 'ce', 'store'])
def test_emulated_multi_gpu_skewed_ranks(emu, exchange):
    """One end of the rank range is made systematically slower (a delay before each of its launches): any operation
    that stores into a peer before that peer has finished reading the buffer shows up as a wrong result."""
    if not FULL and exchange != 'tma':
        pytest.skip('default run: the transfer role only')
    for case, skew in ((MULTI[1], lambda r: 8000*(3 - r)), (MULTI[6], lambda r: 8000*r))[:None if FULL else 1]:
        _multi_case(emu, 4, exchange, case, skew=skew)
    L, _ = emu
    L.sdns_emu_set_skew(0)


@full_only
def test_emulated_multi_gpu_kernel_copy(emu):
    """SDNS_EXCHANGE=kcopy: the send slots are moved by slot_copy_kernel (peer stores) instead of the copy engines."""
    os.environ['SDNS_KCOPY_CTAS'] = '1'
    try:
        _multi_case(emu, 4, 'kcopy', MULTI[1], chunks='2')
    finally:
        os.environ.pop('SDNS_KCOPY_CTAS', None)


@full_only
def test_emulated_multi_gpu_graph_mode_barriers(emu):
    """SDNS_GRAPH=1: from the second identical step on the library 'captures' the step (executed eagerly here) with
    the barrier epochs taken from a device-resident counter and separate flag words; eager operations in between keep
    their own epochs."""
    L, ep = emu
    os.environ['SDNS_GRAPH'], os.environ['SDNS_EXCHANGE'] = '1', 'ce'
    try:
        N, world = (16, 16, 16), 4
        o = so.Oracle(N, dealias='3/2-rule')
        f0 = _state(o, 'NS')
        ref = o.solve(f0, 'NS', 3, 0.002, 0.005)
        u = np.random.RandomState(1).standard_normal((3,)+N)
        uh = o.forward(u)

        def fn(rank, sync):
            L.sdns_emu_set_skew(2000*rank)
            p = ep.EmuPlan(L, N, dealias='3/2-rule', rank=rank, nranks=world)
            p.open_peers(sync(p.handle()))
            k1s, x0s = slice(rank*4, rank*4+4), slice(rank*4, rank*4+4)
            e = [rel_l2(p.rk4(f0[:, :, k1s], 3, 0.002, 0.005), ref[:, :, k1s]),
                 rel_l2(p.backward(uh[:, :, k1s]), u[:, x0s]),
                 rel_l2(p.rk4(f0[:, :, k1s], 3, 0.002, 0.005), ref[:, :, k1s])]
            assert not p.timed_out()
            sync(None)
            p.close()
            return e
        for e in ep.run_ranks(world, fn):
            assert all(x < 1e-11 for x in e), e
    finally:
        os.environ.pop('SDNS_GRAPH', None)
        os.environ.pop('SDNS_EXCHANGE', None)
        L.sdns_emu_set_skew(0)


def test_emulated_multi_gpu_chunk_counts(emu):
    _multi_case(emu, 4, 'ce', MULTI[0], '1')
    _multi_case(emu, 4, 'ce', MULTI[1], '7')


# ---------------------------------------------------------------------------------------------
# the kernels that only exist for long lines (warp-per-line zx with shuffle mirrors, CTA-per-line zy, radix-16
# strided passes, field-parallel F0) on thin grids
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def emu_long():
    import build_emu
    import emu_plan
    os.environ.setdefault('SDNS_EMU_JITTER', '200')
    return emu_plan.load(build_emu.build(lib=os.path.join(build_emu.OUT, 'libsdns_emu_long.so'),
                                         sizes=(8, 12, 256, 384, 512, 768))), emu_plan


@pytest.mark.parametrize('N,precision,dealias', [((8, 8, 256), 'double', '2/3-rule'), ((8, 8, 256), 'single', '3/2-rule'),
                                                 ((8, 8, 512), 'double', '2/3-rule'), ((256, 8, 8), 'double', '2/3-rule'),
                                                 ((8, 256, 8), 'single', '3/2-rule')])
def test_emulated_long_lines(emu_long, N, precision, dealias):
    L, ep = emu_long
    o = so.Oracle(N, precision=precision, dealias=dealias)
    p = ep.EmuPlan(L, N, precision=precision, dealias=dealias)
    rng = np.random.RandomState(7)
    u0 = o.forward(rng.standard_normal((3,)+tuple(N)).astype(o.float)*0.3).astype(o.complex)
    assert rel_l2(p.compute_rhs(u0, 0.01), o.ns_rhs(u0, 0.01)) < TOL[precision]
    assert rel_l2(p.rk4(u0, 1, 0.001, 0.01), o.solve(u0, 'NS', 1, 0.001, 0.01)) < TOL[precision]
    p.close()


@pytest.mark.parametrize('N', [(8, 8, 512), (512, 8, 8)] if FULL else [(512, 8, 8)])
def test_emulated_long_lines_mhd(emu_long, N):
    """TG-MHD 512^3 (BASELINE configs[3]) runs the MHD kernels on 512-point lines: z_kernel<Z_MHD> with CTA barriers (a line
    straddles two warps) and mhd_f0_kernel with its software-pipelined loads."""
    L, ep = emu_long
    o = so.Oracle(N, precision='double', dealias='2/3-rule')
    p = ep.EmuPlan(L, N, precision='double', dealias='2/3-rule', solver='MHD')
    rng = np.random.RandomState(7)
    u0 = o.forward(rng.standard_normal((6,)+tuple(N))*0.3).astype(o.complex)
    assert rel_l2(p.compute_rhs(u0, 0.01, 0.02), o.mhd_rhs(u0, 0.01, 0.02)) < 1e-11
    assert rel_l2(p.rk4(u0, 1, 0.001, 0.01, 0.02), o.solve(u0, 'MHD', 1, 0.001, 0.01, eta=0.02)) < 1e-11


@pytest.fixture(scope='module')
def emu_60():
    import build_emu
    import emu_plan
    return emu_plan.load(build_emu.build(lib=os.path.join(build_emu.OUT, 'libsdns_emu_60.so'), sizes=(8, 12, 60, 90))), emu_plan


@pytest.mark.parametrize('N,precision,dealias,solver', [
    ((60, 8, 8), 'double', '2/3-rule', 'NS'), ((8, 60, 8), 'single', '3/2-rule', 'NS'), ((8, 8, 60), 'double', '3/2-rule', 'NS'),
    ((8, 60, 8), 'double', '3/2-rule', 'VV'), ((8, 8, 60), 'single', '2/3-rule', 'VV'), ((90, 8, 8), 'double', '2/3-rule', 'NS')])
def test_emulated_lengths_60_and_90(emu_60, N, precision, dealias, solver):
    """demo/Isotropic.py's default grid is 60^3 with the 3/2-rule (padded 90): radix-2*2*3*5 / 2*3*3*5 plans with 30
    elements per thread, on each axis, NS and VV Vortex path plus the plain transforms."""
    L, ep = emu_60
    o = so.Oracle(N, precision=precision, dealias=dealias)
    p = ep.EmuPlan(L, N, precision=precision, dealias=dealias, solver=solver)
    f0 = _state(o, solver)
    ref = o.ns_rhs(f0, 0.005) if solver == 'NS' else o.vv_rhs(f0, 0.005)
    assert rel_l2(p.compute_rhs(f0, 0.005), ref) < TOL[precision]
    assert rel_l2(p.rk4(f0, 1, 0.002, 0.005), o.solve(f0, solver, 1, 0.002, 0.005)) < TOL[precision]
    rng = np.random.RandomState(1)
    u = rng.standard_normal((3,)+tuple(N)).astype(o.float)
    assert rel_l2(p.forward(u), o.forward(u)) < TOL[precision]
    assert rel_l2(p.backward(o.forward(u).astype(o.complex)), u) < TOL[precision]
    p.close()


def test_emulated_lengths_60_only_on_the_vortex_path(emu_60):
    L, ep = emu_60
    with pytest.raises(RuntimeError, match='Vortex path only'):
        ep.EmuPlan(L, (60, 8, 8), solver='MHD')
    with pytest.raises(RuntimeError, match='Vortex path only'):
        ep.EmuPlan(L, (8, 8, 60), dealias='3/2-rule', convection='Skewed')


def test_emulated_small_kernels(emu):
    """energy (shuffle + shared-memory reduction), Euler / AB2 steps, cross2, project, add_pressure_diffusion, lincomb / errnorm through the C ABI."""
    import ctypes as C
    L, ep = emu
    vp, dbl = C.c_void_p, C.c_double
    L.sdns_euler_step.argtypes = [vp, vp, vp, dbl, dbl, dbl, vp]
    L.sdns_ab2_step.argtypes = [vp, vp, vp, vp, dbl, C.c_int, dbl, dbl, vp]
    L.sdns_cross2.argtypes = [vp, vp, vp, C.c_int]
    L.sdns_project.argtypes = [vp, vp]
    L.sdns_lincomb.argtypes = [vp, vp, vp, C.c_int, C.POINTER(dbl), C.POINTER(vp), C.c_int]
    L.sdns_errnorm.argtypes = [vp, vp, vp, vp, dbl, dbl, C.c_int, C.POINTER(dbl)]
    N = (16, 16, 16)
    o = so.Oracle(N)
    p = ep.EmuPlan(L, N)
    f0 = _state(o, 'NS')
    e = C.c_double()
    p.chk(L.sdns_energy(p.p, f0.ctypes.data, 3, C.byref(e)))
    assert abs(e.value - o.energy_fourier(f0)) < 1e-12*abs(e.value)
    nu, dt = 0.005, 0.002
    fn = lambda u: o.ns_rhs(u, nu)
    u, rhs = f0.copy(), np.zeros_like(f0)
    ref = f0.copy()
    for _ in range(2):
        p.chk(L.sdns_euler_step(p.p, u.ctypes.data, rhs.ctypes.data, dt, nu, 0.0, None))
        ref = o.forward_euler_step(ref, fn, dt)
    assert rel_l2(u, ref) < 1e-11
    u, u1 = f0.copy(), np.zeros_like(f0)
    ref, r1 = f0.copy(), np.zeros_like(f0)
    for ts in range(3):
        p.chk(L.sdns_ab2_step(p.p, u.ctypes.data, u1.ctypes.data, rhs.ctypes.data, dt, ts, nu, 0.0, None))
        ref, r1 = o.ab2_step(ref, r1, fn, dt, ts)
    assert rel_l2(u, ref) < 1e-11
    c = np.zeros_like(f0)
    p.chk(L.sdns_cross2(p.p, c.ctypes.data, f0.ctypes.data, 0))
    assert rel_l2(c, o.cross2(o.K, f0)) < 1e-13
    v = f0.copy() + 0.1*(np.random.RandomState(2).standard_normal(f0.shape) + 0j)
    w = v.copy()
    p.chk(L.sdns_project(p.p, w.ctypes.data))
    pref = v - np.sum(o.K_over_K2*v, 0)*np.array([np.broadcast_to(k, o.sshape) for k in o.K])
    assert rel_l2(w, pref) < 1e-13
    # add_pressure_diffusion_NS on its own (cython_solvers.in:40-80)
    L.sdns_add_pressure_diffusion.argtypes = [vp, vp, vp, dbl, vp]
    du, ph = v.copy(), np.zeros(o.sshape, dtype=o.complex)
    p.chk(L.sdns_add_pressure_diffusion(p.p, du.ctypes.data, f0.ctypes.data, 0.0123, ph.ctypes.data))
    du_ref, p_ref = o.add_pressure_diffusion(v.copy(), f0, 0.0123)
    assert rel_l2(du, du_ref) < 1e-14 and rel_l2(ph, p_ref) < 1e-14
    # diagnostics and forcing of demo/Isotropic.py on the device (Isotropic.py:88-118, 161-184, 243-247)
    i32 = C.c_int
    L.sdns_energy_weighted.argtypes = [vp, vp, i32, vp, i32, C.POINTER(dbl)]
    L.sdns_scale_field.argtypes = [vp, vp, i32, vp, i32, dbl, dbl]
    L.sdns_set_mode.argtypes = [vp, vp, i32, i32, i32, i32, dbl, dbl]
    L.sdns_enstrophy.argtypes = [vp, C.c_void_p, C.POINTER(dbl)]
    L.sdns_divergence_norm.argtypes = [vp, vp, C.POINTER(dbl)]
    L.sdns_spectrum.argtypes = [vp, vp, i32, i32, C.POINTER(dbl), C.POINTER(dbl)]
    g = v.copy()
    g[:, 0, 0, 0] = 1.5 - 0.5j
    k2_mask = np.where(o.K2 <= 3**2, 1, 0)
    ref, e_new, e_low, alpha = o.forcing_rescale(g.copy(), 3, 1.1*o.energy_fourier(v))
    p.chk(L.sdns_set_mode(p.p, g.ctypes.data, 3, 0, 0, 0, 0.0, 0.0))
    assert np.all(g[:, 0, 0, 0] == 0)
    wgt = np.ascontiguousarray(k2_mask, dtype=np.float64)
    p.chk(L.sdns_energy_weighted(p.p, g.ctypes.data, 3, wgt.ctypes.data, 1, C.byref(e)))
    assert abs(e.value - e_low) < 1e-12*abs(e_low)
    p.chk(L.sdns_energy_weighted(p.p, g.ctypes.data, 3, None, 1, C.byref(e)))
    assert abs(e.value - o.energy_fourier(g)) < 1e-12*abs(e.value)
    fac = np.ascontiguousarray(alpha*k2_mask + (1-k2_mask), dtype=np.float64)
    g_aff = g.copy()
    p.chk(L.sdns_scale_field(p.p, g.ctypes.data, 3, fac.ctypes.data, 1, 1.0, 0.0))
    assert rel_l2(g, ref) < 1e-15
    p.chk(L.sdns_scale_field(p.p, g_aff.ctypes.data, 3, wgt.ctypes.data, 1, float(alpha), 1.0))      # from the mask itself
    assert rel_l2(g_aff, ref) < 1e-15
    fac32 = fac.astype(np.float32)
    g2 = v.copy()
    p.chk(L.sdns_scale_field(p.p, g2.ctypes.data, 3, fac32.ctypes.data, 0, 1.0, 0.0))
    assert rel_l2(g2, v*fac32) < 1e-15
    p.chk(L.sdns_enstrophy(p.p, v.ctypes.data, C.byref(e)))
    assert abs(e.value - o.enstrophy(v)) < 1e-12*abs(e.value)
    # Parseval needs the spectrum of a real field without Nyquist modes (what the solver's mask_nyquist leaves)
    w_h = (o.forward(np.random.RandomState(3).standard_normal((3,)+tuple(N)))*o.mask).astype(o.complex)
    p.chk(L.sdns_divergence_norm(p.p, w_h.ctypes.data, C.byref(e)))
    assert abs(e.value - o.divergence_norm(w_h)) < 1e-11*abs(e.value)
    Ek_ref, bins = o.spectrum(v)
    nb = len(bins)
    sums, cnts = (dbl*nb)(), (dbl*nb)()
    p.chk(L.sdns_spectrum(p.p, v.ctypes.data, 3, nb, sums, cnts))
    sums, cnts = np.array(sums[:]), np.array(cnts[:])
    Ek = np.zeros(nb)
    for i in range(nb-1):
        if cnts[i]:
            Ek[i] = ((bins[i+1])**3 - bins[i]**3)*sums[i]*(4./3.*np.pi)/cnts[i]
    assert np.allclose(Ek, Ek_ref, rtol=1e-12, atol=0) and cnts.sum() > 0
    out = np.zeros_like(f0)
    coeffs = (dbl*2)(0.25, -1.5)
    arrs = (vp*2)(f0.ctypes.data, c.ctypes.data)
    p.chk(L.sdns_lincomb(p.p, out.ctypes.data, v.ctypes.data, 2, coeffs, arrs, 3))
    assert rel_l2(out, v + 0.25*f0 - 1.5*c) < 1e-14
    en = (dbl*3)()
    p.chk(L.sdns_errnorm(p.p, f0.ctypes.data, v.ctypes.data, c.ctypes.data, 1e-6, 1e-3, 3, en))
    sc = 1e-6 + np.maximum(np.abs(f0), np.abs(v))*1e-3
    assert np.allclose(np.array(en[:]), np.sum(np.abs(c/sc)**2, axis=(1, 2, 3)), rtol=1e-11)
    p.close()


# ---- doubly periodic solvers (csrc/sdns2d_api.cu) ----------------------------------------------------------------
import sdns_oracle2d as so2      # noqa: E402


class Emu2D(object):
    """numpy front end of the sdns2d_* entry points of the emulated library ("device" pointers are numpy buffers)."""
    def __init__(self, L, N, Lbox=(2*np.pi,)*2, precision='double', dealias='2/3-rule', solver='NS2D', mask_nyquist=True):
        import ctypes as C
        from spectraldns_b200 import _lib
        self.C, self.L = C, _lib.bind2d(L)
        cfg = _lib.Sdns2dConfig()
        cfg.abi_version = 1
        for i in range(2):
            cfg.N[i], cfg.L[i], cfg.kcut[i] = int(N[i]), float(Lbox[i]), -1
        cfg.precision = 1 if precision == 'double' else 0
        cfg.dealias = _lib.DEALIAS[dealias]
        cfg.solver = _lib.SOLVER2D[solver]
        cfg.mask_nyquist = int(mask_nyquist)
        self.p = C.c_void_p()
        self.chk(L.sdns2d_plan_create(C.byref(self.p), C.byref(cfg)))
        n = C.c_size_t()
        self.chk(L.sdns2d_workspace_bytes(self.p, C.byref(n)))
        self._ws = np.full(n.value + 512, 0xFF, dtype=np.uint8)
        base = self._ws.ctypes.data
        self.chk(L.sdns2d_plan_set_workspace(self.p, C.c_void_p(base + (-base) % 256), n.value))
        sp, ph, pd = (C.c_int32*2)(), (C.c_int32*2)(), (C.c_int32*2)()
        self.chk(L.sdns2d_shapes(self.p, C.byref(sp), C.byref(ph), C.byref(pd)))
        self.sshape, self.pshape, self.dshape = tuple(sp), tuple(ph), tuple(pd)
        self.real = np.float64 if precision == 'double' else np.float32
        self.cplx = np.complex128 if precision == 'double' else np.complex64

    def chk(self, rc):
        if rc:
            raise RuntimeError(self.L.sdns2d_last_error().decode())

    def forward(self, u, padded=False):
        u = np.ascontiguousarray(u, dtype=self.real)
        out = np.full(u.shape[:-2]+self.sshape, np.nan, dtype=self.cplx)
        self.chk(self.L.sdns2d_forward(self.p, int(padded), u.shape[0], u.ctypes.data, out.ctypes.data))
        return out

    def backward(self, uh, padded=False):
        uh = np.ascontiguousarray(uh, dtype=self.cplx)
        out = np.full(uh.shape[:-2]+(self.dshape if padded else self.pshape), np.nan, dtype=self.real)
        self.chk(self.L.sdns2d_backward(self.p, int(padded), uh.shape[0], uh.ctypes.data, out.ctypes.data))
        return out

    def compute_rhs(self, uh, nu, Ri=0.0, Pr=1.0, source=None, want_p=False):
        uh = np.ascontiguousarray(uh, dtype=self.cplx)
        rhs = np.full(uh.shape, np.nan, dtype=self.cplx)
        ph = np.full(self.sshape, np.nan, dtype=self.cplx)
        self.chk(self.L.sdns2d_compute_rhs(self.p, rhs.ctypes.data, uh.ctypes.data, nu, Ri, Pr,
                                           source.ctypes.data if source is not None else None, ph.ctypes.data if want_p else None))
        return (rhs, ph) if want_p else rhs

    def rk4(self, uh, nsteps, dt, nu, Ri=0.0, Pr=1.0):
        u = np.ascontiguousarray(uh, dtype=self.cplx).copy()
        u1, u2 = np.full(u.shape, np.nan, dtype=self.cplx), np.full(u.shape, np.nan, dtype=self.cplx)
        for _ in range(nsteps):
            self.chk(self.L.sdns2d_rk4_step(self.p, u.ctypes.data, u1.ctypes.data, u2.ctypes.data, dt, nu, Ri, Pr, None))
        return u

    def close(self):
        self.L.sdns2d_plan_destroy(self.p)


def _state2d(o, ncomp, seed=5):
    rng = np.random.RandomState(seed)
    U = rng.standard_normal((ncomp,)+o.N)
    X = o.mesh()
    U[0] += 2*np.sin(X[0])*np.cos(X[1])
    U[1] -= 2*np.sin(X[1])*np.cos(X[0])
    uh = o.forward(U.astype(o.float))
    if o.mask is not None:
        uh = uh*o.mask
    return uh.astype(o.complex)


@pytest.mark.parametrize('solver,N,Lbox,dealias,precision', [
    ('NS2D', (16, 16), (2*np.pi, 2*np.pi), '2/3-rule', 'double'), ('NS2D', (32, 16), (6*np.pi, 4*np.pi), '3/2-rule', 'double'),
    ('NS2D', (16, 48), (2*np.pi, 2*np.pi), 'None', 'single'), ('NS2D', (8, 32), (2*np.pi, 4*np.pi), '2/3-rule', 'single'),
    ('Bq2D', (16, 16), (2*np.pi, 2*np.pi), '2/3-rule', 'double'), ('Bq2D', (16, 32), (6*np.pi, 4*np.pi), '3/2-rule', 'double'),
    ('Bq2D', (32, 16), (2*np.pi, 2*np.pi), '3/2-rule', 'single')])
def test_emulated_2d_solvers(emu, solver, N, Lbox, dealias, precision):
    """NS2D (solvers/NS2D.py:13-51) and Bq2D (solvers/Bq2D.py:101-186): transforms on T and Tp, ComputeRHS with the
    pressure, two RK4 steps, Forward Euler / AB2 and the stand-alone 2-D operators against oracle/sdns_oracle2d.py."""
    import ctypes as C
    L, _ = emu
    o = so2.Oracle2D(N, L=Lbox, precision=precision, dealias=dealias)
    p = Emu2D(L, N, Lbox, precision, dealias, solver)
    tol = TOL[precision]
    nc = 3 if solver == 'Bq2D' else 2
    rng = np.random.RandomState(2)
    u = rng.standard_normal((nc,)+tuple(N)).astype(o.float)
    assert rel_l2(p.forward(u), o.forward(u)) < tol
    assert rel_l2(p.backward(o.forward(u).astype(o.complex)), u) < tol
    f0 = _state2d(o, nc)
    assert rel_l2(p.backward(f0, padded=True), o._bwd_p(f0)) < tol
    up = rng.standard_normal((nc,)+tuple(o.M)).astype(o.float)
    assert rel_l2(p.forward(up, padded=True), o._fwd_p(up)) < tol
    nu, Ri, Pr, dt = 0.01, 0.1, 0.7, 0.01
    if solver == 'NS2D':
        ref, pref = o.ns2d_rhs(f0, nu, return_p=True)
        fn = lambda v: o.ns2d_rhs(v, nu)
    else:
        ref, pref = o.bq2d_rhs(f0, nu, Ri, Pr, return_p=True)
        fn = lambda v: o.bq2d_rhs(v, nu, Ri, Pr)
    rhs, ph = p.compute_rhs(f0, nu, Ri, Pr, want_p=True)
    assert rel_l2(rhs, ref) < tol and rel_l2(ph, pref) < tol
    assert rel_l2(p.rk4(f0, 2, dt, nu, Ri, Pr), o.solve(f0, solver, 2, dt, nu, Ri, Pr)) < tol
    # ForwardEuler and AB2 (maths/integrators.py:161-175)
    uu, rr = f0.copy(), np.zeros_like(f0)
    p.chk(p.L.sdns2d_euler_step(p.p, uu.ctypes.data, rr.ctypes.data, dt, nu, Ri, Pr, None))
    assert rel_l2(uu, o.forward_euler_step(f0, fn, dt)) < tol
    uu, u1 = f0.copy(), np.zeros_like(f0)
    refu, r1 = f0.copy(), np.zeros_like(f0)
    for ts in range(3):
        p.chk(p.L.sdns2d_ab2_step(p.p, uu.ctypes.data, u1.ctypes.data, rr.ctypes.data, dt, ts, nu, Ri, Pr, None))
        refu, r1 = o.ab2_step(refu, r1, fn, dt, ts)
    assert rel_l2(uu, refu) < tol
    # stand-alone operators
    c = np.zeros(o.sshape, dtype=o.complex)
    p.chk(p.L.sdns2d_cross2(p.p, c.ctypes.data, f0.ctypes.data))
    assert rel_l2(c, o.cross2(f0[:2])) < (1e-14 if precision == 'double' else 1e-6)
    du = (f0[::-1]*0.3).astype(o.complex).copy()
    ph = np.zeros(o.sshape, dtype=o.complex)
    p.chk(p.L.sdns2d_add_pressure_diffusion(p.p, du.ctypes.data, f0.ctypes.data, nu, Ri, Pr, ph.ctypes.data))
    if solver == 'NS2D':
        dref, pr = o.add_pressure_diffusion_ns2d((f0[::-1]*0.3).astype(o.complex), f0, nu)
    else:
        dref, pr = o.add_pressure_diffusion_bq2d((f0[::-1]*0.3).astype(o.complex), f0, nu, Ri, Pr)
    assert rel_l2(du, dref) < (1e-14 if precision == 'double' else 1e-6) and rel_l2(ph, pr) < (1e-14 if precision == 'double' else 1e-6)
    if solver == 'NS2D':
        src = (0.1*f0[::-1]).astype(o.complex).copy()
        assert rel_l2(p.compute_rhs(f0, nu, source=src), o.ns2d_rhs(f0, nu, source=src)) < tol
    p.close()


def test_emulated_2d_taylor_green_known_answer(emu):
    """tests/TG2D.py:41-52 through the C ABI: kinetic energy of the 2-D Taylor-Green vortex after 20 RK4 steps against
    the analytic decay, to the reference's ntol = 7 digits."""
    L, _ = emu
    N = (32, 32)
    o = so2.Oracle2D(N)
    p = Emu2D(L, N)
    nu, dt, nsteps = 0.01, 0.05, 20
    u = p.rk4(so2.taylor_green_2d(o), nsteps, dt, nu)
    U = p.backward(u)
    k = np.sum(U.astype(np.float64)**2)/np.prod(N)/2
    ke = 0.25*np.exp(-2*nu*nsteps*dt)**2
    assert round(float(k - ke), 7) == 0
    p.close()
