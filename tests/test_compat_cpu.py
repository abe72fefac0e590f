"""CPU tests of the reference-facing host layer (no GPU): config semantics, module surface,
library symbols.  Compute calls are covered by the -m gpu tests."""
import ctypes
import os
import re
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def compat():
    sys.path.insert(0, ROOT)
    from spectraldns_b200 import run
    run.activate()
    import spectralDNS
    return spectralDNS


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every function include/sdns_b200.h declares."""
    hdr = open(os.path.join(ROOT, 'include', 'sdns_b200.h')).read()
    declared = set(re.findall(r'^\s*(?:int|const char\*)\s+(sdns(?:2d)?_\w+)\s*\(', hdr, flags=re.M))
    assert len(declared) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, 'spectraldns_b200', 'libsdns_b200.so'))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from spectraldns_b200 import _lib
    assert set(_lib.SYMBOLS) | set(_lib.SYMBOLS2D) == declared
    assert lib.sdns_abi_version() == 1
    for n in (8, 12, 16, 24, 256, 384, 768, 1024, 2048, 3072):
        assert lib.sdns_size_supported(n, 1) == 0
    assert lib.sdns_size_supported(60, 1) == 0 and lib.sdns_size_supported(90, 0) == 0      # demo/Isotropic.py default grid
    assert lib.sdns_size_supported(100, 1) != 0 and lib.sdns_size_supported(36, 1) != 0


def test_plan_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from spectraldns_b200 import _lib
    from spectraldns_b200.plan import Plan
    with pytest.raises(_lib.SdnsError):
        Plan((32, 32, 32))
    # and straight through the C ABI
    L = _lib.lib()
    cfg = _lib.SdnsConfig()
    cfg.abi_version = 1
    for i in range(3):
        cfg.N[i], cfg.L[i], cfg.kcut[i] = 32, 2*np.pi, -1
    cfg.precision, cfg.dealias, cfg.nranks = 1, 1, 1
    p = ctypes.c_void_p()
    rc = L.sdns_plan_create(ctypes.byref(p), ctypes.byref(cfg))
    assert rc == -3 and b'no CPU fallback' in L.sdns_last_error()


def test_config_matches_reference_semantics(compat):
    """Same option names/defaults/derived values as the reference's config.py (:96-239)."""
    config = compat.config
    config.update({'nu': 0.000625, 'dt': 0.01, 'T': 0.1, 'L': [2*np.pi, '4*pi', 6*np.pi], 'M': [4, 5, 6]})
    ns = config.triplyperiodic.parse_args(['--precision', 'single', 'MHD', '--eta', '0.02'])
    config.params.update(vars(ns))
    P = config.params
    assert list(P.N) == [16, 32, 64] and P.M.flags.writeable is False
    assert np.allclose(P.L, [2*np.pi, 4*np.pi, 6*np.pi])
    assert np.allclose(P.dx, P.L/P.N)
    assert isinstance(P.nu, np.float32) and isinstance(P.dt, np.float32) and isinstance(P.eta, np.float32)
    assert P.solver == 'MHD' and P.dealias == '2/3-rule' and P.integrator == 'RK4'
    assert P.convection == 'Vortex' and P.decomposition == 'slab' and P.mask_nyquist is True
    assert P.write_result == 1e8 and P.checkpoint == 1e8 and P.ntol == 7 and P.optimization == ''
    ns = config.triplyperiodic.parse_args(['--no-mask_nyquist', '--dealias', '3/2-rule', '--optimization',
                                           'cython', '--N', ] if False else ['--no-mask_nyquist', 'NS'])
    assert ns.mask_nyquist is False and ns.solver == 'NS'
    # demos add their own arguments and N trumps M (tests/TG.py:139-142)
    config.triplyperiodic.add_argument('--N', default=[32, 32, 32], nargs=3)
    ns = config.triplyperiodic.parse_args(['NS'])
    config.params.update(vars(ns))
    assert list(config.params.N) == [32, 32, 32]
    with pytest.raises(SystemExit):
        config.triplyperiodic.parse_args(['--dealias', 'bogus', 'NS'])
    config.update({'L': [2*np.pi]*3, 'M': [6, 6, 6]})


def test_solver_module_surface(compat):
    """The names solve(), the demos and the tests read from a solver module."""
    import importlib
    for name in ('NS', 'VV', 'MHD'):
        m = importlib.import_module('spectralDNS.solvers.' + name)
        for attr in ('get_context', 'getConvection', 'ComputeRHS', 'getintegrator', 'comm', 'rank',
                     'num_processes', 'params', 'profiler', 'Timer', 'MemoryUsage', 'create_profile',
                     'cross1', 'cross2', 'project', 'update', 'regression_test', 'additional_callback',
                     'end_of_tstep', 'set_source', 'solve_linear', 'datatypes', 'get_divergence',
                     'work_arrays', 'HDF5File', 'optimizer', 'device_state'):
            assert hasattr(m, attr), (name, attr)
    from spectralDNS.solvers import NS, VV
    for attr in ('get_velocity', 'get_curl', 'get_pressure', 'set_velocity', 'add_pressure_diffusion'):
        assert hasattr(NS, attr)
    for attr in ('get_velocity', 'get_curl', 'compute_velocity', 'add_linear'):
        assert hasattr(VV, attr)
    assert NS.datatypes('single')[:2] == (np.float32, np.complex64)
    import shenfun
    for attr in ('FunctionSpace', 'TensorProductSpace', 'VectorSpace', 'CompositeSpace', 'Array',
                 'Function', 'CachedArrayDict', 'ShenfunFile'):
        assert hasattr(shenfun, attr)
    from shenfun.fourier import energy_fourier  # noqa: F401
    from mpi4py_fft import generate_xdmf        # noqa: F401


def test_hdf5file_cadence_and_killfile(compat, tmp_path, monkeypatch):
    """h5io/HDF5File.py:64-120: results every write_result steps, checkpoint every `checkpoint`."""
    from spectraldns_b200.io import HDF5File
    monkeypatch.chdir(tmp_path)
    u_hat, u = np.ones((3, 4, 4, 3), dtype=complex), np.ones((3, 4, 4, 4))
    f = HDF5File('run', checkpoint={'space': None, 'data': {'0': {'U': [u_hat]}}},
                 results={'space': None, 'data': {'U': [u]}})
    P = compat.config.AttributeDict(tstep=1, t=0.01, write_result=2, checkpoint=4, filemode='w')
    f.update(P)
    assert f.cfile is None and f.wfile is None
    P.tstep = 2
    f.update(P)
    assert os.path.exists('run_w_t2.npz') and not os.path.exists('run_c.npz')      # one results archive per written step
    P.tstep, P.t = 4, 0.04
    f.update(P)
    z = np.load('run_c.npz')
    assert int(z['attr__tstep']) == 4 and abs(float(z['attr__t']) - 0.04) < 1e-15 and 'U/3D/0' in z.files
    open('killspectraldns', 'w').close()
    P.tstep = 5
    with pytest.raises(SystemExit):
        f.update(P)
    assert not os.path.exists('killspectraldns')
    f.close()
    # restart from the checkpoint (h5io/HDF5File.py init_from_file): global shape + local slice travel with the data
    from spectraldns_b200.io import read_global
    fields, attrs = read_global('run_c', 'U/3D/')
    assert attrs['tstep'] == 5 and np.array_equal(fields['U/3D/0'], u_hat)


_KILL_WORKER = r'''
import os, sys
sys.path[:0] = [%(root)r, os.path.join(%(root)r, 'spectraldns_b200', 'compat')]
import numpy as np
import torch.distributed as dist
dist.init_process_group('gloo')
rank = dist.get_rank()
os.chdir(%(cwd)r)
from spectraldns_b200.io import HDF5File, read_global
from spectralDNS import config


class Space(object):           # a slab-distributed spectral space: axis 1 split over the ranks
    def global_shape(self, spectral=True):
        return (4, 4, 3)
    def local_slice(self, spectral=True):
        return (slice(0, 4), slice(2*rank, 2*rank + 2), slice(0, 3))


u_hat = np.full((3, 4, 2, 3), rank + 1, dtype=complex)
f = HDF5File('run', checkpoint={'space': Space(), 'data': {'0': {'U': [u_hat]}}}, results={'space': None, 'data': {}})
P = config.AttributeDict(tstep=1, t=0.01, write_result=10**8, checkpoint=10**8, filemode='w')
f.update(P)
dist.barrier()
if rank == 1:                  # only ONE rank sees the file appear before the next update
    open('killspectraldns', 'w').close()
dist.barrier()
P.tstep = 2
try:
    f.update(P)
    print('rank %%d NOT STOPPED' %% rank)
except SystemExit:
    print('rank %%d stopped' %% rank)
dist.barrier()
if rank == 0:
    fields, attrs = read_global('run_c', 'U/3D/')
    g = fields['U/3D/0']
    assert g.shape == (3, 4, 4, 3) and (g[:, :, :2] == 1).all() and (g[:, :, 2:] == 2).all() and attrs['tstep'] == 2
    assert not os.path.exists('killspectraldns')
    print('KILL_WORKER_OK')
dist.destroy_process_group()
'''


def test_killfile_is_collective_over_ranks(tmp_path):
    """Two gloo ranks: the kill file is visible to one rank only when update() runs; both must checkpoint and stop
    (the reference all-reduces `found`, h5io/HDF5File.py), rank 0 removes the file, and the per-rank checkpoint
    archives reassemble into the global array."""
    import subprocess
    script = tmp_path / 'kill_worker.py'
    script.write_text(_KILL_WORKER % {'root': ROOT, 'cwd': str(tmp_path)})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29611', str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and 'KILL_WORKER_OK' in r.stdout and r.stdout.count('stopped') == 2 and 'NOT STOPPED' not in r.stdout, r.stdout[-3000:]


def test_timer_interface(compat, capsys):
    from spectralDNS.utilities import Timer
    t = Timer()
    t(); t()
    t.final(True)
    out = capsys.readouterr().out
    assert 'Fastest = (' in out and 'Slowest = (' in out and 'Time = ' in out


def test_c_abi_error_behaviour():
    """Argument validation happens before any CUDA call, so it is testable without a GPU: every entry
    returns a negative sdns_status and sdns_last_error() names the problem."""
    from spectraldns_b200 import _lib
    L = _lib.lib()

    def cfg(**kw):
        c = _lib.SdnsConfig()
        c.abi_version = 1
        for i in range(3):
            c.N[i], c.L[i], c.kcut[i] = 32, 2*np.pi, -1
        c.precision, c.dealias, c.solver, c.convection, c.nranks = 1, 1, 0, 0, 1
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    p = ctypes.c_void_p()
    assert L.sdns_plan_create(ctypes.byref(p), None) == -1
    assert L.sdns_plan_create(ctypes.byref(p), ctypes.byref(cfg(abi_version=99))) == -1
    assert b'ABI' in L.sdns_last_error()
    assert L.sdns_plan_create(ctypes.byref(p), ctypes.byref(cfg(precision=7))) == -1
    assert L.sdns_plan_create(ctypes.byref(p), ctypes.byref(cfg(nranks=9))) == -1
    assert L.sdns_plan_create(ctypes.byref(p), ctypes.byref(cfg(nranks=2, rank=0, decomposition=1))) == -1
    assert b'slab' in L.sdns_last_error()
    assert L.sdns_plan_create(ctypes.byref(p), ctypes.byref(cfg(solver=1, convection=1))) == -1      # VV + Divergence
    assert b'VV' in L.sdns_last_error()
    assert L.sdns_plan_create(ctypes.byref(p), ctypes.byref(cfg(solver=2, convection=0))) == -1      # MHD + Vortex
    assert L.sdns_plan_create(ctypes.byref(p), ctypes.byref(cfg(convection=7))) == -1
    # null plans
    assert L.sdns_sync(None) == -1 and L.sdns_forward(None, 0, 1, None, None) != 0
    n = ctypes.c_size_t()
    assert L.sdns_workspace_bytes(None, ctypes.byref(n)) == -1


def test_slab_and_plan_size_rules_agree():
    """spectraldns_b200.slab, the numpy replay of the exchange that the gloo tests drive, covers the even split only and
    says so; the library itself also takes extents the rank count does not divide (N // P per rank, the first N % P one
    more: tests/test_kernels_emulated.py::test_emulated_multi_gpu_uneven_slabs)."""
    from spectraldns_b200.slab import SlabLayout
    for N, P, ok in (((64, 64, 64), 8, True), ((64, 36, 64), 8, False), ((36, 64, 64), 8, False),
                     ((48, 48, 48), 3, True), ((32, 32, 32), 5, False)):
        try:
            SlabLayout(N, P, 0)
            good = True
        except ValueError:
            good = False
        assert good == ok, (N, P)


def test_checkpoint_metadata_carries_the_slice_step(tmp_path, monkeypatch):
    """Cyclic axis-1 ownership (SDNS_K1_LAYOUT=cyclic): a rank's block is the strided slice [r::P].  The archives record
    start AND step per axis, read_global reassembles them, and archives without a step (contiguous blocks) still read."""
    monkeypatch.chdir(tmp_path)
    from spectraldns_b200.io import ShenfunFile, read_global
    g = (np.arange(3*4*6*3).reshape(3, 4, 6, 3) + 1j).astype(complex)

    class Space(object):
        def __init__(self, r):
            self.r = r

        def global_shape(self, spectral=True):
            return (4, 6, 3)

        def local_slice(self, spectral=True):
            return (slice(0, 4), slice(self.r, 6, 2), slice(0, 3))
    for r in range(2):
        gshape, start, step = ShenfunFile('unused', Space(r))._meta(g[:, :, r::2])       # what the writer records
        assert gshape == (3, 4, 6, 3) and start == (0, 0, r, 0) and step == (1, 1, 2, 1)
        m = np.array([3, 4, 6, 3, 0, 0, r, 0, 1, 1, 2, 1], dtype=np.int64)
        np.savez('cyc_c_rank%d.npz' % r, **{'U/3D/0': g[:, :, r::2], 'meta__U/3D/0': m, 'attr__tstep': np.array(7)})
    fields, attrs = read_global('cyc_c', 'U/3D/')
    assert np.array_equal(fields['U/3D/0'], g) and attrs['tstep'] == 7
    for r in range(2):                  # the older layout of the metadata: shape + start, contiguous blocks
        m = np.array([3, 4, 6, 3, 0, 0, 3*r, 0], dtype=np.int64)
        np.savez('blk_c_rank%d.npz' % r, **{'U/3D/0': g[:, :, 3*r:3*r + 3], 'meta__U/3D/0': m})
    fields, _ = read_global('blk_c', 'U/3D/')
    assert np.array_equal(fields['U/3D/0'], g)


def test_convection_building_blocks_of_the_solver_modules(compat):
    """NS.Cross / standard_convection / divergence_convection and MHD.set_Elsasser / divergenceConvection (reference
    NS.py:131-162, MHD.py:89-110) exist in the host mirror for user code that calls them; on host arrays and a numpy
    space they reproduce the oracle's restatement of the reference's convection forms."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import sdns_oracle as so
    from spectralDNS.solvers import NS, MHD
    N = (12, 8, 16)
    o = so.Oracle(N, dealias='None')

    class Space(object):                         # T == Tp without dealiasing
        def forward(self, u, out=None):
            r = o.forward(u[None])[0]
            if out is not None:
                out[...] = r
                return out
            return r

        def backward(self, uh, out=None):
            r = o.backward(np.asarray(uh)[None])[0]
            if out is not None:
                out[...] = r
                return out
            return r

    class VSpace(Space):
        def forward(self, u, out=None):
            r = o.forward(u)
            if out is not None:
                out[...] = r
                return out
            return r
    T = Space()
    u_hat = so.isotropic_field(o, seed=5)
    u = o.backward(u_hat)
    K = o.K
    rel = lambda a, b: np.linalg.norm((a - b).ravel())/np.linalg.norm(b.ravel())
    got = NS.standard_convection(np.zeros_like(u_hat), u, u_hat, None, T, K)
    assert rel(-got, o.ns_conv(u_hat, 'Standard')) < 1e-12
    got = NS.divergence_convection(np.zeros_like(u_hat), u, None, T, K)
    assert rel(-got, o.ns_conv(u_hat, 'Divergence')) < 1e-12
    skew = NS.divergence_convection(NS.standard_convection(np.zeros_like(u_hat), u, u_hat, None, T, K), u, None, T, K, add=True)
    assert rel(-0.5*skew, o.ns_conv(u_hat, 'Skewed')) < 1e-12
    curl = o.backward(o.cross2(K, u_hat))
    assert rel(NS.Cross(np.zeros_like(u_hat), u, curl, None, VSpace()), o.ns_conv(u_hat, 'Vortex')) < 1e-12
    # MHD: the convection part of the oracle's right-hand side (no Nyquist mask, nu = eta = 0, pressure projected out)
    ub_hat = so.isotropic_field(o, seed=6, ncomp=6)
    ub = o.backward(ub_hat)
    c = MHD.divergenceConvection(np.zeros_like(ub_hat), ub[:3] + ub[3:], ub[:3] - ub[3:], T, K,
                                 np.zeros((3, 3) + u_hat.shape[1:], dtype=complex))
    P_hat = np.sum(c[:3]*o.K_over_K2, 0)
    for i in range(3):
        c[i] -= P_hat*K[i]
    o.mask = None
    assert rel(c, o.mhd_rhs(ub_hat, 0.0, 0.0)) < 1e-12


def test_integrator_functions_and_submodule_paths(compat):
    """spectralDNS.maths.integrators.RK4 / ForwardEuler / AB2 (reference maths/integrators.py:150-175) on host arrays with
    a solver object whose ComputeRHS is a numpy function, against the oracle's restatement; the submodule import paths
    of the reference (maths.cross, maths.maths, maths.integrators) and utilities.cleanup / inheritdocstrings exist."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import sdns_oracle as so
    from spectralDNS.maths.integrators import RK4, ForwardEuler, AB2, getintegrator          # noqa: F401
    from spectralDNS.maths.cross import cross1, cross2                                       # noqa: F401
    from spectralDNS.maths.maths import project                                              # noqa: F401
    from spectralDNS.utilities import cleanup, inheritdocstrings
    o = so.Oracle((8, 8, 8))
    lam = -0.7 + 0.3j

    class Solver(object):
        @staticmethod
        def ComputeRHS(rhs, u, solver, **ctx):
            rhs[...] = lam*u + 0.1*u*u
            return rhs
    fn = lambda u: lam*u + 0.1*u*u
    rng = np.random.RandomState(2)
    u0 = (rng.standard_normal((3, 4, 4, 3)) + 1j*rng.standard_normal((3, 4, 4, 3)))
    a, b, dt = np.array([1./6., 1./3., 1./3., 1./6.]), np.array([0.5, 0.5, 1.]), 0.05
    u = u0.copy()
    got, _, _ = RK4(u, np.zeros_like(u), np.zeros_like(u), np.zeros_like(u), a, b, dt, Solver, {})
    assert np.allclose(got, o.rk4_step(u0.copy(), fn, dt), rtol=1e-14, atol=0)
    u = u0.copy()
    got, _, _ = ForwardEuler(u, np.zeros_like(u), dt, Solver, {})
    assert np.allclose(got, o.forward_euler_step(u0.copy(), fn, dt), rtol=1e-14, atol=0)
    u, u1 = u0.copy(), np.zeros_like(u0)
    ref, r1 = u0.copy(), np.zeros_like(u0)
    for tstep in range(3):
        u, _, _ = AB2(u, u1, np.zeros_like(u), dt, tstep, Solver, {})
        ref, r1 = o.ab2_step(ref, r1, fn, dt, tstep)
    assert np.allclose(u, ref, rtol=1e-13, atol=0)

    class A(object):
        def f(self):
            """doc of A.f"""

    @inheritdocstrings
    class B(A):
        def f(self):
            pass
    assert B.f.__doc__ == 'doc of A.f' and callable(cleanup)
