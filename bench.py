#!/usr/bin/env python
"""bench.py -- time per RK4 step and grid-points*steps/s of the pseudo-spectral NS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--grid 512] [--precision double|single] [--dealias 2/3-rule|3/2-rule]
                    [--solver NS|VV|MHD]

Default workload: Taylor-Green NS 512^3 double, RK4, 2/3-rule on 1 B200 -- the first size BASELINE.json's metric is
quoted on (512^3 / 1024^3 / 2048^3); weak scaling doubles one axis per doubling of the GPU count, so N = 8 is the
1024^3 north-star configuration.  `--grid 256` is BASELINE configs[1].
One "step" = one RK4 step (4 right-hand sides = 36 scalar 3-D FFTs + fused pointwise work).
Rank 0 prints ONE JSON line.  Before the timed region every rank runs a small parity gate (64^3 NS and MHD, one
right-hand side and two RK4 steps on a broadband field) against the CPU oracle and the run fails if it is off.
--impl reference times the reference's OWN solver modules (solvers/NS.py, maths/integrators.py, its Cython kernels
compiled into oracle/_ref, all staged under baseline/_ref by oracle/stage_reference_scripts.sh) over the numpy
stand-ins for the absent shenfun / mpi4py-fft (oracle/shim, pocketfft with all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import spectraldns_b200  # noqa: E402,F401  (sets CUDA_DEVICE_MAX_CONNECTIONS before the CUDA context exists)

NU, DT = 0.000625, 0.01       # tests/TG.py:131-133 of the reference (32^3)


def dt_for(a, N):
    """The reference's dt = 0.01 where RK4 is stable with it, else the largest stable one with a margin: the advection term
    has eigenvalues up to i*sum_i kc_i*max|v_i| (kc_i = N_i/3 kept modes; |v| <= 1 for the Taylor-Green velocity, 2 for the
    Elsasser fields of TG-MHD), and RK4 is stable on the imaginary axis below 2.83.  dt = 0.01 on 1024^3 (MHD: 512^3) grows
    round-off in the highest modes by a factor > 3 (> 50) per step.  The work per step does not depend on dt."""
    vmax = 2.0 if a.solver == 'MHD' else 1.0
    return min(DT, 1.4/(vmax*sum(n/3.0 for n in N)))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=('ours', 'reference'))
    ap.add_argument('--grid', type=int, default=512)
    ap.add_argument('--precision', default='double', choices=('single', 'double'))
    ap.add_argument('--dealias', default='2/3-rule', choices=('2/3-rule', '3/2-rule', 'None'))
    ap.add_argument('--solver', default='NS', choices=('NS', 'VV', 'MHD'))
    ap.add_argument('--scaling', default='weak', choices=('weak', 'strong'),
                    help='N>1: weak = grid grows with the GPU count (per-GPU work fixed), strong = fixed grid')
    ap.add_argument('--k1-layout', default='auto', choices=('auto', 'blocks', 'cyclic'),
                    help='N>1: which axis-1 modes a rank owns (include/sdns_b200.h).  auto = cyclic on 3 or more GPUs under the '
                         '2/3 rule (every rank keeps the same number of modes), else the reference\'s contiguous blocks')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    ap.add_argument('--no-parity', action='store_true', help='skip the parity gate and the field-level check at size')
    ap.add_argument('--timeline', default=None,
                    help='write PATH.rank<r>.json: start/end of every kernel, peer copy and barrier of one RK4 step')
    return ap.parse_args()


def grid_for(a, world):
    """N=1: grid^3.  N>1 weak scaling: double N0, N1, N2 in turn (256^3 -> 512x256x256 -> 512x512x256
    -> 512^3 on 8 GPUs) so that every GPU keeps grid^3 points; strong scaling keeps grid^3."""
    N = [a.grid]*3
    if world > 1 and a.scaling == 'weak':
        w, i = world, 0
        while w > 1:
            N[i % 3] *= 2
            w //= 2
            i += 1
    return tuple(N)


def k1_layout_for(a, world):
    if a.k1_layout != 'auto':
        return a.k1_layout if world > 1 else 'blocks'
    return 'cyclic' if (world >= 3 and a.dealias == '2/3-rule') else 'blocks'


def workload_name(a, N=None):
    N = N or (a.grid,)*3
    return 'Taylor-Green %s %dx%dx%d %s RK4 %s slab' % (a.solver, N[0], N[1], N[2], a.precision, a.dealias)


def peaks():
    f = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(f):
        try:
            return float(json.load(open(f))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------
# The reference's own CPU implementation of the path (cpu_baseline leg and --impl reference arm)
# ---------------------------------------------------------------------------------------------
def model_bytes(a, N):
    """SURVEY.md section 8(d) three-pass model: bytes one RK4 step moves when every scalar 3-D transform is three
    axis passes that read and write their array once and only the stage update costs extra (276 F for NS / VV with
    the 2/3 rule, 487.5 F with 3/2 padding, 480 F / 832.5 F for MHD; F = one complex field at the unpadded size)."""
    F = float(N[0])*N[1]*(N[2]//2+1)*(16 if a.precision == 'double' else 8)
    per_transform = 11.875 if a.dealias == '3/2-rule' else 6.0
    ntr, nst = (15, 30.0) if a.solver == 'MHD' else (9, 15.0)
    return 4*(ntr*per_transform + nst)*F


def _reference_modules():
    """Import the UNMODIFIED reference package staged under baseline/_ref over oracle/shim.  Returns None when it
    has not been staged (then the callers fall back to the numpy port in oracle/sdns_oracle.py)."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.exists(os.path.join(ref, 'spectralDNS', 'solvers', 'NS.py')):
        return None
    for p in (ref, os.path.join(ROOT, 'oracle', 'shim')):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import spectralDNS as ref_pkg
        from spectralDNS import config, get_solver, solve
    except Exception as e:                                  # pragma: no cover
        sys.stderr.write('reference package not importable: %r\n' % (e,))
        return None
    if 'baseline' not in os.path.abspath(ref_pkg.__file__):
        return None                                         # the drop-in `spectralDNS` got in first: not the reference
    return config, get_solver, solve


def reference_steps(a, Ns, nsteps, warm, want_state=False):
    """Run the reference solver (its get_solver / get_context / solve loop, --optimization cython) for warm + nsteps
    RK4 steps of the Taylor-Green problem on an Ns^3-type grid; per-step wall times from its own Timer hook.
    Returns (seconds per step over the timed steps, fastest step, kind, final spectral state or None)."""
    import contextlib
    import io
    import numpy as np
    mods = _reference_modules()
    M = [int(round(np.log2(n))) for n in Ns]
    if mods is None or any(2**m != n for m, n in zip(M, Ns)):
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import sdns_oracle as so
        o = so.Oracle(Ns, precision=a.precision, dealias=a.dealias)
        u = o.forward(so.taylor_green_mhd(o) if a.solver == 'MHD' else so.taylor_green(o))
        if a.solver == 'VV':
            u = o.cross2(o.K, u)
        ts = []
        for _ in range(warm + nsteps):
            t0 = time.perf_counter()
            u = o.solve(u, a.solver, 1, dt_for(a, Ns), NU, eta=0.01)
            ts.append(time.perf_counter() - t0)
        ts = ts[warm:]
        return sum(ts)/len(ts), min(ts), 'port', (u if want_state else None)
    config, get_solver, solve = mods
    config.update({'nu': NU, 'dt': dt_for(a, Ns), 'T': dt_for(a, Ns)*(warm + nsteps), 'eta': 0.01,
                   'convection': 'Divergence' if a.solver == 'MHD' else 'Vortex'})
    with contextlib.redirect_stdout(io.StringIO()):
        solver = get_solver(parse_args=['--M'] + [str(m) for m in M] +
                            ['--precision', a.precision, '--dealias', a.dealias, '--optimization', 'cython',
                             '--integrator', 'RK4', a.solver])
        ctx = solver.get_context()
    X = ctx.X
    U = ctx.UB if a.solver == 'MHD' else ctx.U
    U[:] = 0
    U[0] = np.sin(X[0])*np.cos(X[1])*np.cos(X[2])
    U[1] = -np.cos(X[0])*np.sin(X[1])*np.cos(X[2])
    if a.solver == 'MHD':
        U[3] = np.sin(X[0])*np.sin(X[1])*np.cos(X[2])
        U[4] = np.cos(X[0])*np.cos(X[1])*np.cos(X[2])
    uh = ctx.UB_hat if a.solver == 'MHD' else ctx.U_hat
    space = ctx.VM if a.solver == 'MHD' else ctx.VT
    uh[:] = space.forward(U, uh) if hasattr(space, 'forward') else uh
    if a.solver == 'VV':
        ctx.W_hat = solver.cross2(ctx.W_hat, ctx.K, ctx.U_hat)
    stamps = []

    class StepTimer(solver.Timer):                          # the reference's own per-step hook (utilities/__init__.py:33-39)
        def __call__(self):
            super().__call__()
            stamps.append(time.perf_counter())
    solver.Timer = StepTimer
    config.params.t = 0.0
    config.params.tstep = 0
    t_start = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        solve(solver, ctx)
    assert len(stamps) == warm + nsteps, (len(stamps), warm, nsteps)
    ts = [b - a_ for a_, b in zip([t_start] + stamps[:-1], stamps)][warm:]
    state = None
    if want_state:
        state = np.array(ctx.W_hat if a.solver == 'VV' else uh)
    return sum(ts)/len(ts), min(ts), 'reference', state


def sample_grid(a, N, budget_s, nsteps):
    """Largest grid N / 2^j whose (nsteps) reference steps fit the time budget, from a 64^3-type probe step and
    N log N scaling (the CPU sample of a workload too large to step on the host within the bench's time box)."""
    import numpy as np
    probe = tuple(max(32, n//(N[0]//64)) if N[0] > 64 else n for n in N)
    t, _, _, _ = reference_steps(a, probe, 1, 1)
    pts = lambda g: float(g[0])*g[1]*g[2]*np.log2(float(g[0])*g[1]*g[2])
    g = tuple(N)
    while g[0] > probe[0] and 1.6*t*pts(g)/pts(probe)*nsteps > budget_s:     # 1.6: large grids fall out of cache
        g = tuple(n//2 for n in g)
    return g


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    N = grid_for(a, max(1, a.gpus))
    warm, steps = max(a.warmup, 0), max(a.steps, 1)
    Ns = sample_grid(a, N, 150.0, warm + steps)
    spt, fastest, kind, _ = reference_steps(a, Ns, steps, warm)
    pts = float(Ns[0])*Ns[1]*Ns[2]
    val = pts/spt
    note = ('the reference\'s own solvers/%s.py + maths/integrators.py RK4 + its Cython kernels (--optimization cython) over the '
            'numpy stand-ins for shenfun / mpi4py-fft (scipy.fft pocketfft, all host threads, single rank); the reference\'s '
            'FFTW / MPI stack is absent from the image' % a.solver) if kind == 'reference' else (
            'numpy port of the reference path (oracle/sdns_oracle.py): the staged reference package was not found')
    line = {
        'impl': 'reference', 'metric': 'grid_points_steps_per_s', 'value': val, 'unit': 'points*steps/s',
        'n_gpus': a.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': spt*1e3,
        'higher_is_better': True, 'scaling': a.scaling if a.gpus > 1 else 'weak', 'vs_baseline': None,
        'dtype': 'f64' if a.precision == 'double' else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(a, N), 'grid': list(N), 'integrator': 'RK4', 'note': note},
        'cpu_baseline': {'value': val, 'unit': 'points*steps/s', 'cores': os.cpu_count(), 'kind': kind,
                         'fastest_step_ms': fastest*1e3,
                         'sample': ('%d timed + %d warm-up full RK4 steps of the same problem on a %dx%dx%d grid' % ((steps, warm) + tuple(Ns))) +
                                   ('' if tuple(Ns) == tuple(N) else ' (the %dx%dx%d workload itself does not fit the time box on the '
                                    'host; points*steps/s is size-normalised)' % tuple(N))},
        'e2e': {'value': val, 'unit': 'points*steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# parity: the CPU oracle as the checker of the GPU path, inside the run the driver makes
# ---------------------------------------------------------------------------------------------
def parity_gate(rank, world, local, layout='blocks'):
    """64^3 NS (double and single) and MHD: one right-hand side and two RK4 steps on a seeded broadband field, every
    rank against its slab of the single-process oracle.  Returns {case: rel L2}, the largest ratio err / tol."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import sdns_oracle as so
    from spectraldns_b200.plan import Plan
    out, worst = {}, 0.0
    for N, prec, dealias, solver in (((64, 64, 64), 'double', '2/3-rule', 'NS'), ((64, 64, 64), 'double', '2/3-rule', 'MHD'),
                                     ((64, 64, 64), 'single', '3/2-rule', 'NS')):
        tol = 1e-11 if prec == 'double' else 1e-4
        o = so.Oracle(N, precision=prec, dealias=dealias)
        p = Plan(N, precision=prec, dealias=dealias, solver=solver, device=local, rank=rank, nranks=world, k1_layout=layout)
        nc = 6 if solver == 'MHD' else 3
        f0 = so.isotropic_field(o, seed=3, ncomp=nc).astype(o.complex)
        k1s = p.k1_slice
        nu, eta, dt = 0.005, 0.01, 0.002
        r_ref = o.ns_rhs(f0, nu) if solver == 'NS' else o.mhd_rhs(f0, nu, eta)
        d_u = p.to_device(f0[:, :, k1s])
        rhs = p.to_host(p.compute_rhs(p.empty_spectral(), d_u, nu, eta))
        u1, u2 = p.empty_spectral(), p.empty_spectral()
        for _ in range(2):
            p.rk4_step(d_u, u1, u2, dt, nu, eta)
        s_ref = o.solve(f0, solver, 2, dt, nu, eta=eta)
        rel = lambda x, y: float(np.linalg.norm((x.astype(np.complex128) - y).ravel())/np.linalg.norm(y.ravel()))
        e = max(rel(rhs, r_ref[:, :, k1s]), rel(p.to_host(d_u), s_ref[:, :, k1s]))
        if p.comm_timed_out():
            e = float('inf')
        out['%s_%d_%s_%s' % (solver, N[0], prec, dealias)] = e
        worst = max(worst, e/tol)
        del p
    return out, worst


def state_parity(a, Ns, ref_state, nsteps):
    """The same Taylor-Green run on the GPU path (host field -> forward -> nsteps RK4 steps) against the reference
    solver's final spectral state: relative L2 over the whole field."""
    import numpy as np
    from spectraldns_b200.plan import Plan
    p = Plan(Ns, precision=a.precision, dealias=a.dealias, solver=a.solver)
    X = np.meshgrid(*[np.arange(n)*2*np.pi/n for n in Ns], indexing='ij')
    U = np.zeros((p.ncomp,) + tuple(Ns), dtype=p.float)
    U[0] = np.sin(X[0])*np.cos(X[1])*np.cos(X[2])
    U[1] = -np.cos(X[0])*np.sin(X[1])*np.cos(X[2])
    if a.solver == 'MHD':
        U[3] = np.sin(X[0])*np.sin(X[1])*np.cos(X[2])
        U[4] = np.cos(X[0])*np.cos(X[1])*np.cos(X[2])
    u = p.forward(p.to_device(U))
    if a.solver == 'VV':
        u = p.cross2(p.empty_spectral(), u)
    u1, u2 = p.empty_spectral(), p.empty_spectral()
    for _ in range(nsteps):
        p.rk4_step(u, u1, u2, dt_for(a, Ns), NU, 0.01)
    got = p.to_host(u).astype(np.complex128)
    err = float(np.linalg.norm((got - ref_state).ravel())/np.linalg.norm(ref_state.ravel()))
    tol = 1e-11 if a.precision == 'double' else 1e-4
    return {'grid': list(Ns), 'steps': nsteps, 'rel_l2_vs_reference_solver': err, 'tol': tol, 'ok': bool(err < tol)}


# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '50'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                out['sm_max_mhz'] = float(r[1])
                for nme, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(nme)
            except Exception:
                pass
        if sm:
            sm.sort()
            out['sm_mhz'] = sm[len(sm)//2]
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        return out


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from spectraldns_b200.plan import Plan

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product has no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if not a.no_parity:
        cases, worst = parity_gate(rank, world, local, k1_layout_for(a, world))
        t = torch.tensor([worst], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parity = {'rel_l2': cases, 'tol': {'double': 1e-11, 'single': 1e-4}, 'worst_err_over_tol_all_ranks': float(t.item()),
                  'what': 'GPU path vs CPU oracle (oracle/sdns_oracle.py), seeded broadband field, 1 RHS + 2 RK4 steps, '
                          'every rank its slab; rank 0 values shown'}
        if not float(t.item()) < 1.0:
            if rank == 0:
                print(json.dumps({'error': 'parity gate failed', 'parity': parity}))
            raise SystemExit(3)

    N = grid_for(a, world)
    p = Plan(N, precision=a.precision, dealias=a.dealias, solver=a.solver, device=local,
             rank=rank, nranks=world, k1_layout=k1_layout_for(a, world))
    # synthetic Taylor-Green field generated on the device (tests/TG.py:23-28 / tests/TGMHD.py:4-12)
    X = [torch.arange(n, dtype=torch.float64, device='cuda')*2*np.pi/n for n in N]
    X[0] = X[0][p.x0_slice]                      # this rank's slab of physical space
    s0, c0 = torch.sin(X[0])[:, None, None], torch.cos(X[0])[:, None, None]
    s1, c1 = torch.sin(X[1])[None, :, None], torch.cos(X[1])[None, :, None]
    c2 = torch.cos(X[2])[None, None, :]
    U = p.empty_physical()
    U[0] = (s0*c1*c2).to(p.tfloat)
    U[1] = (-c0*s1*c2).to(p.tfloat)
    if a.solver == 'MHD':
        U[3] = (s0*s1*c2).to(p.tfloat)
        U[4] = (c0*c1*c2).to(p.tfloat)
    u = p.forward(U)
    del U
    if a.solver == 'VV':
        w = p.cross2(p.empty_spectral(), u)
        u = w
    u1, u2 = p.empty_spectral(), p.empty_spectral()
    eta, dt = 0.01, dt_for(a, N)
    state_bytes = u.numel()*u.element_size()

    for _ in range(max(a.warmup, 3)):
        p.rk4_step(u, u1, u2, dt, NU, eta)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = p.launch_count()
    xb0, xf0 = p.xfer_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        p.rk4_step(u, u1, u2, dt, NU, eta)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)/a.steps
    launches = p.launch_count() - l0
    xb1, xf1 = p.xfer_stats()
    xfer_bytes_step, xfer_flushes_step = (xb1 - xb0)/a.steps, (xf1 - xf0)/a.steps
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    et = torch.tensor([p.energy(u)/2], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(et)
    energy = float(et.item())
    assert np.isfinite(energy) and 0 < energy < 1, energy
    assert not p.comm_timed_out()

    # ---- end to end through the host-buffer call (H2D + step + D2H every step) ------------
    host = torch.empty(u.shape, dtype=u.dtype, pin_memory=True)
    host.copy_(u)
    ne = max(3, min(a.steps, 10))
    p.rk4_steps_host(host, u, u1, u2, 1, dt, NU, eta)       # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(ne):
        p.rk4_steps_host(host, u, u1, u2, 1, dt, NU, eta)
    torch.cuda.synchronize()
    te = (time.perf_counter() - t0)/ne
    t = torch.tensor([te], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    te = float(t.item())

    # ---- per-kernel device times (separate pass, CUDA events around every launch) ----------
    p.profile(True)
    npf = max(2, min(a.steps, 5))
    for _ in range(npf):
        p.rk4_step(u, u1, u2, dt, NU, eta)
    prof = p.profile_read()
    copies = p.profile_read_copies()
    p.profile(False)
    if a.timeline:
        barrier()
        p.profile(True, timeline=True)
        p.rk4_step(u, u1, u2, dt, NU, eta)
        rows = p.profile_timeline()
        p.profile(False)
        with open('%s.rank%d.json' % (a.timeline, rank), 'w') as f:
            json.dump([{'what': w, 't0_ms': t0, 't1_ms': t1, 'bytes': b} for w, t0, t1, b in rows], f)
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pts = float(N[0])*N[1]*N[2]
    value = pts/(ms*1e-3)
    peak, peak_src = peaks()
    tot = sum(v[0] for v in prof.values())
    kern = max(prof, key=lambda k: prof[k][0])
    kms, kn, kb = prof[kern][:3]
    achieved = kb/kms*1e-6 if kms > 0 else 0.0       # bytes/ms -> GB/s
    traffic = None          # DRAM bytes need an ncu capture: the per-round captures are summarised under profiles/
    step_bytes = sum(v[2] for v in prof.values())/npf
    mbytes = model_bytes(a, N)/world
    line = {
        'metric': 'grid_points_steps_per_s', 'value': value, 'unit': 'points*steps/s',
        'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3), 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': a.scaling if world > 1 else 'weak', 'vs_baseline': None,
        'dtype': 'f64' if a.precision == 'double' else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(a, N), 'grid': list(N), 'integrator': 'RK4',
                   'transforms_per_step': 60 if a.solver == 'MHD' else 36,
                   'multi_gpu': ('slab decomposition over %d GPUs, %s; %s scaling' % (world, (
                       'send slots moved over NVLink peer memory by a bulk-async transfer role inside the FFT pass kernels '
                       '(no NCCL on the data path)' if xfer_bytes_step > 0 else
                       'exchange in chunks by the copy engines over NVLink peer memory underneath the FFT passes '
                       '(no NCCL on the data path)' if copies[2] else
                       'transposes are peer-memory stores fused into the FFT passes (no NCCL on the data path)'),
                       a.scaling)) if world > 1 else 'single GPU',
                   'k1_layout': (p.k1_layout + (' (rank r owns the axis-1 modes r, r + P, ...: every rank keeps the same number of '
                                                'modes under the 2/3 rule; same global field as the reference\'s contiguous blocks)'
                                                if p.k1_layout == 'cyclic' else ' (the reference\'s contiguous axis-1 blocks)')) if world > 1 else None,
                   'l2': 'inputs larger than L2 (state %.0f MB, scratch %.0f MB)' % (state_bytes/1e6, p.workspace_bytes/1e6),
                   'timing': 'CUDA events on the launch stream, max over ranks',
                   'dt': dt, 'nu': NU,
                   'kinetic_energy_after_run': energy},
        'clocks': clocks,
        'e2e': {'value': pts/te, 'unit': 'points*steps/s', 'ms_per_step': te*1e3,
                'h2d_bytes_per_step': state_bytes*world, 'd2h_bytes_per_step': state_bytes*world,
                'call': 'sdns_rk4_steps_host: pinned host state -> device, one RK4 step, device -> host'},
        'gpu_launches': launches,
        'parity': parity,
        'roofline': {'bound': 'hbm', 'kernel': kern, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved/peak, 'traffic': traffic, 'peak_source': peak_src,
                     'kernel_share_of_step': kms/tot if tot else None,
                     'kernel_ms_per_launch': kms/kn, 'algorithmic_bytes_per_launch': kb/kn,
                     'step_algorithmic_GB': step_bytes*1e-9,
                     'step_achieved_GBps': step_bytes*1e-9/(ms*1e-3),
                     'step_frac': step_bytes*1e-9/(ms*1e-3)/peak,
                     'survey_8d_model_GB_per_gpu': mbytes*1e-9,
                     'survey_8d_model_frac': mbytes*1e-9/(ms*1e-3)/peak,
                     'all_kernels': {k: {'ms_per_launch': v[0]/v[1], 'launches_per_step': v[1]/npf,
                                         'GBps': v[2]/v[0]*1e-6, 'share': v[0]/tot} for k, v in prof.items()}},
    }
    if world > 1 and xfer_bytes_step > 0:
        # NVLink side of the roofline (rank 0's view), transfer-role exchange: the bytes this rank sent per step over
        # the whole step time = the SUSTAINED rate per direction
        sus = xfer_bytes_step*1e-9/(ms*1e-3)
        line['nvlink'] = {'bytes_per_step_per_gpu': xfer_bytes_step, 'sustained_GBps_per_direction': sus,
                          'peak_measured_GBps': 770.0, 'peak_nominal_GBps': 900.0, 'frac': sus/770.0,
                          'frac_of_nominal': sus/900.0, 'transfer_only_launches_per_step': xfer_flushes_step,
                          'exchange': os.environ.get('SDNS_EXCHANGE', 'tma (default for 3 or more GPUs)'),
                          'note': 'send slots moved by bulk-async (TMA) copies issued from a transfer role inside the FFT pass '
                                  'kernels (csrc/xfer.cuh); frac = sustained over the whole step / 770 GB/s measured peer copy'}
    elif world > 1 and copies[2]:
        # NVLink side of the roofline (rank 0's view), copy-engine exchange: bytes sent per step; rate while the
        # busiest per-peer copy stream is busy, and sustained over the whole step
        xbytes, xms = copies[1]/npf, copies[0]/npf
        line['nvlink'] = {'bytes_per_step_per_gpu': xbytes, 'copies_per_step': copies[2]/npf,
                          'copy_stream_busy_ms_per_step': xms,
                          'while_busy_GBps_per_direction': xbytes*1e-9/(xms*1e-3) if xms else None,
                          'sustained_GBps_per_direction': xbytes*1e-9/(ms*1e-3),
                          'peak_measured_GBps': 770.0, 'peak_nominal_GBps': 900.0,
                          'frac': xbytes*1e-9/(ms*1e-3)/770.0,
                          'while_busy_frac_of_measured': (xbytes*1e-9/(xms*1e-3))/770.0 if xms else None,
                          'exchange': 'ce',
                          'note': 'strided cudaMemcpy2DAsync per peer and chunk on per-peer streams, concurrent with '
                                  'the FFT passes; frac = bytes sent per step / step time / 770 GB/s measured peer copy '
                                  '(sustained); while_busy = rate while the busiest copy stream is busy'}
    elif world > 1:
        # NVLink side of the roofline (rank 0's view): bytes the exchange passes store into peers
        xk = {k: v for k, v in prof.items() if v[3] > 0}
        xbytes = sum(v[3] for v in xk.values())/npf
        xms = sum(v[0] for v in xk.values())/npf
        line['nvlink'] = {'bytes_per_step_per_gpu': xbytes, 'exchange_kernels': sorted(xk),
                          'exchange_kernel_ms_per_step': xms,
                          'achieved_GBps_per_direction': xbytes*1e-9/(xms*1e-3) if xms else None,
                          'peak_measured_GBps': 770.0, 'peak_nominal_GBps': 900.0,
                          'frac_of_measured': (xbytes*1e-9/(xms*1e-3))/770.0 if xms else None,
                          'note': 'transposes are peer stores issued by the FFT pass in front of them, so the '
                                  'kernel time also covers that pass\'s local HBM traffic'}
    if world == 1 and not a.no_cpu_baseline:
        # the reference's own solver on the host cores, bounded: 2 timed RK4 steps on the largest grid N / 2^j that
        # fits ~cpu_seconds; its final state is also the field-level parity reference for the GPU path at that size
        Ns = sample_grid(a, N, a.cpu_seconds, 3)
        spt, fastest, kind, ref_state = reference_steps(a, Ns, 2, 1, want_state=not a.no_parity)
        spts = float(Ns[0])*Ns[1]*Ns[2]
        line['cpu_baseline'] = {'value': spts/spt, 'unit': 'points*steps/s', 'cores': os.cpu_count(), 'kind': kind,
                                'ms_per_step': spt*1e3, 'fastest_step_ms': fastest*1e3,
                                'sample': '2 timed + 1 warm-up full RK4 steps of the same Taylor-Green problem on a %dx%dx%d grid '
                                          'with the reference\'s own solver modules + Cython kernels over scipy.fft (all host '
                                          'threads); points*steps/s is size-normalised' % tuple(Ns)}
        if ref_state is not None:
            del p, u, u1, u2
            torch.cuda.empty_cache()
            line['parity']['at_size'] = state_parity(a, Ns, ref_state, 3)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
