#!/usr/bin/env python
"""bench.py -- time per RK4 step and grid-points*steps/s of the pseudo-spectral NS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--grid 256] [--precision double|single] [--dealias 2/3-rule|3/2-rule]
                    [--solver NS|VV|MHD]

Default workload = BASELINE.json configs[1]: Taylor-Green NS 256^3 double, RK4, 2/3-rule, 1 B200.
One "step" = one RK4 step (4 right-hand sides = 36 scalar 3-D FFTs + fused pointwise work).
Rank 0 prints ONE JSON line.  --impl reference times the CPU oracle port (numpy + scipy.fft with
all host threads; the reference's own stack -- shenfun/mpi4py/pyfftw -- is absent from the image).
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

NU, DT = 0.000625, 0.01       # tests/TG.py:131-133 of the reference


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=('ours', 'reference'))
    ap.add_argument('--grid', type=int, default=256)
    ap.add_argument('--precision', default='double', choices=('single', 'double'))
    ap.add_argument('--dealias', default='2/3-rule', choices=('2/3-rule', '3/2-rule', 'None'))
    ap.add_argument('--solver', default='NS', choices=('NS', 'VV', 'MHD'))
    ap.add_argument('--scaling', default='weak', choices=('weak', 'strong'),
                    help='N>1: weak = grid grows with the GPU count (per-GPU work fixed), strong = fixed grid')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
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
# CPU oracle port (also the --impl reference arm)
# ---------------------------------------------------------------------------------------------
def cpu_oracle_steps(a, N, max_steps, max_seconds, warm=0):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import numpy as np
    import sdns_oracle as so
    o = so.Oracle(N, precision=a.precision, dealias=a.dealias)
    if a.solver == 'MHD':
        u = o.forward(so.taylor_green_mhd(o))
    else:
        u = o.forward(so.taylor_green(o))
        if a.solver == 'VV':
            u = o.cross2(o.K, u)
    eta = 0.01
    for _ in range(warm):
        u = o.solve(u, a.solver, 1, DT, NU, eta=eta)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        u = o.solve(u, a.solver, 1, DT, NU, eta=eta)
        n += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = (time.perf_counter() - t0)/n
    assert np.isfinite(u).all()
    return dt, n, os.cpu_count()


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # bounded: stop after ~150 s of timed work whatever K is
    warm = 1 if a.warmup > 0 else 0
    N = grid_for(a, max(1, a.gpus))
    spt, n, cores = cpu_oracle_steps(a, N, a.steps, 150.0, warm=warm)
    pts = float(N[0])*N[1]*N[2]
    val = pts/spt
    line = {
        'impl': 'reference', 'metric': 'grid_points_steps_per_s', 'value': val, 'unit': 'points*steps/s',
        'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': spt*1e3,
        'higher_is_better': True, 'scaling': a.scaling if a.gpus > 1 else 'weak', 'vs_baseline': None,
        'dtype': 'f64' if a.precision == 'double' else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(a, N), 'grid': list(N), 'integrator': 'RK4',
                   'note': 'CPU oracle port of the reference path (numpy + scipy.fft pocketfft, workers=all cores, '
                           'single rank); the reference stack shenfun/mpi4py-fft/pyfftw/mpirun is absent from the image'},
        'cpu_baseline': {'value': val, 'unit': 'points*steps/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d full RK4 steps of the %dx%dx%d workload (%d warm-up)' % (n, N[0], N[1], N[2], warm)},
        'e2e': {'value': val, 'unit': 'points*steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


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

    N = grid_for(a, world)
    p = Plan(N, precision=a.precision, dealias=a.dealias, solver=a.solver, device=local,
             rank=rank, nranks=world)
    # synthetic Taylor-Green field generated on the device (tests/TG.py:23-28 / tests/TGMHD.py:4-12)
    X = [torch.arange(n, dtype=torch.float64, device='cuda')*2*np.pi/n for n in N]
    M0l = N[0]//world
    X[0] = X[0][rank*M0l:(rank+1)*M0l]          # this rank's slab of physical space
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
    eta = 0.01
    state_bytes = u.numel()*u.element_size()

    for _ in range(max(a.warmup, 3)):
        p.rk4_step(u, u1, u2, DT, NU, eta)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = p.launch_count()
    xb0, xf0 = p.xfer_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        p.rk4_step(u, u1, u2, DT, NU, eta)
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
    p.rk4_steps_host(host, u, u1, u2, 1, DT, NU, eta)       # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(ne):
        p.rk4_steps_host(host, u, u1, u2, 1, DT, NU, eta)
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
        p.rk4_step(u, u1, u2, DT, NU, eta)
    prof = p.profile_read()
    copies = p.profile_read_copies()
    p.profile(False)
    if a.timeline:
        barrier()
        p.profile(True, timeline=True)
        p.rk4_step(u, u1, u2, DT, NU, eta)
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
    traffic = None
    tf = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get('%s_%d_%s' % (kern, a.grid, a.precision))
        except Exception:
            traffic = None
    step_bytes = sum(v[2] for v in prof.values())/npf
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
                   'l2': 'inputs larger than L2 (state %.0f MB, scratch %.0f MB)' % (state_bytes/1e6, p.workspace_bytes/1e6),
                   'timing': 'CUDA events on the launch stream, max over ranks',
                   'kinetic_energy_after_run': energy},
        'clocks': clocks,
        'e2e': {'value': pts/te, 'unit': 'points*steps/s', 'ms_per_step': te*1e3,
                'h2d_bytes_per_step': state_bytes*world, 'd2h_bytes_per_step': state_bytes*world,
                'call': 'sdns_rk4_steps_host: pinned host state -> device, one RK4 step, device -> host'},
        'gpu_launches': launches,
        'roofline': {'bound': 'hbm', 'kernel': kern, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved/peak, 'traffic': traffic, 'peak_source': peak_src,
                     'kernel_share_of_step': kms/tot if tot else None,
                     'kernel_ms_per_launch': kms/kn, 'algorithmic_bytes_per_launch': kb/kn,
                     'step_algorithmic_GB': step_bytes*1e-9,
                     'step_achieved_GBps': step_bytes*1e-9/(ms*1e-3),
                     'step_frac': step_bytes*1e-9/(ms*1e-3)/peak,
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
                          'exchange': os.environ.get('SDNS_EXCHANGE', 'tma'),
                          'note': 'send slots moved by bulk-async (TMA) copies issued from a transfer role inside the FFT pass '
                                  'kernels (csrc/xfer.cuh); frac = sustained over the whole step / 770 GB/s measured peer copy'}
    elif world > 1 and copies[2]:
        # NVLink side of the roofline (rank 0's view), copy-engine exchange: bytes sent per step; rate while the
        # busiest per-peer copy stream is busy, and sustained over the whole step
        xbytes, xms = copies[1]/npf, copies[0]/npf
        line['nvlink'] = {'bytes_per_step_per_gpu': xbytes, 'copies_per_step': copies[2]/npf,
                          'copy_stream_busy_ms_per_step': xms,
                          'achieved_GBps_per_direction': xbytes*1e-9/(xms*1e-3) if xms else None,
                          'sustained_over_step_GBps': xbytes*1e-9/(ms*1e-3),
                          'peak_measured_GBps': 770.0, 'peak_nominal_GBps': 900.0,
                          'frac_of_measured': (xbytes*1e-9/(xms*1e-3))/770.0 if xms else None,
                          'note': 'strided cudaMemcpy2DAsync per peer and chunk on per-peer streams, concurrent with '
                                  'the FFT passes; busy time = summed copy durations of the busiest stream'}
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
        spt, n, cores = cpu_oracle_steps(a, N, 50, a.cpu_seconds)
        line['cpu_baseline'] = {'value': pts/spt, 'unit': 'points*steps/s', 'cores': cores, 'kind': 'port',
                                'ms_per_step': spt*1e3,
                                'sample': '%d full RK4 steps of the same %d^3 workload with the numpy/scipy.fft '
                                          'oracle (workers=all cores)' % (n, a.grid)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
