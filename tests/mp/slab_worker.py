"""Worker for tests/test_gpu_multi.py: one process per GPU (torchrun), slab decomposition.
Every rank builds the same seeded global field, keeps its slab, runs the B200 path and compares
its slab of the result with the slab of the single-process CPU oracle."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]

import torch
import torch.distributed as dist
import sdns_oracle as so
from spectraldns_b200.plan import Plan


def rel(a, b):
    return float(np.linalg.norm((a.astype(np.complex128)-b).ravel())/max(np.linalg.norm(b.ravel()), 1e-300))


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    fails = []
    # SLAB_K1_LAYOUT=cyclic: rank r owns the axis-1 modes [r::P] instead of the reference's contiguous block
    layout = os.environ.get('SLAB_K1_LAYOUT', 'blocks')
    cases = [((32, 32, 32), 'double', '2/3-rule', 'NS'), ((32, 32, 32), 'double', '3/2-rule', 'NS'),
             ((64, 32, 16), 'double', '2/3-rule', 'VV'), ((32, 32, 32), 'single', '2/3-rule', 'NS'),
             ((16, 32, 64), 'double', '2/3-rule', 'MHD'), ((32, 16, 32), 'double', 'None', 'NS'),
             ((128, 128, 128), 'double', '2/3-rule', 'NS'),
             # low axis-1 cutoff: with 4 or more ranks the middle ones own no mode that survives the truncation
             ((32, 32, 32), 'double', '2/3-rule', 'NS', (-1, 3, -1)), ((32, 64, 32), 'single', '2/3-rule', 'MHD', (-1, 5, -1)),
             # the other convection forms of NS.getConvection (NS.py:164-201), VV with padding, MHD in single precision
             ((32, 32, 32), 'double', '2/3-rule', 'NS', None, 'Standard'), ((32, 32, 16), 'double', '3/2-rule', 'NS', None, 'Divergence'),
             ((16, 32, 32), 'double', '2/3-rule', 'NS', None, 'Skewed'), ((32, 32, 32), 'single', '3/2-rule', 'NS', None, 'Skewed'),
             ((32, 32, 32), 'double', '3/2-rule', 'VV'), ((32, 32, 32), 'single', '3/2-rule', 'MHD'),
             # extents the rank count does not divide (N // P per rank, the first N % P ranks one more): axis 1 on 8 ranks;
             # axis 0 on 4 ranks and both on 8
             ((24, 12, 16), 'double', '2/3-rule', 'NS'), ((90, 60, 16), 'double', '2/3-rule', 'NS')]
    if os.environ.get('SLAB_CASES'):            # a subset, by index (the 8-GPU box is paid by the second)
        cases = [cases[int(i)] for i in os.environ['SLAB_CASES'].split(',')]
    for case in cases:
        N, prec, dealias, solver = case[:4]
        kcut = case[4] if len(case) > 4 else None
        conv = case[5] if len(case) > 5 else None
        tol = 1e-11 if prec == 'double' else 1e-4
        if layout == 'cyclic' and N[1] % world:
            continue                                    # the cyclic ownership needs N[1] divisible by the ranks
        o = so.Oracle(N, precision=prec, dealias=dealias, kcut=kcut)
        p = Plan(N, precision=prec, dealias=dealias, solver=solver, device=local, rank=rank, nranks=world, kcut=kcut,
                 convection=conv, k1_layout=layout)
        N1l = N[1]//world
        k1s = p.k1_slice
        if N[1] % world == 0:
            assert k1s == (slice(rank, N[1], world) if layout == 'cyclic' else slice(rank*N1l, (rank+1)*N1l)), k1s
        x0s, x0ps = p.x0_slice, p.x0p_slice
        M0l, Mp0l = x0s.stop - x0s.start, x0ps.stop - x0ps.start
        if N[1] % world == 0:
            assert p.spectral_shape == (N[0], N1l, N[2]//2+1), p.spectral_shape
        assert p.physical_shape == (M0l, N[1], N[2]) and p.padded_shape == (Mp0l, o.M[1], o.M[2])
        nc = 6 if solver == 'MHD' else 3
        rng = np.random.RandomState(11)
        u = rng.standard_normal((nc,)+tuple(N)).astype(o.float)
        # plain transforms T.forward / T.backward on slabs
        uh = p.to_host(p.forward(p.to_device(u[:, x0s])))
        ref = o.forward(u)
        e1 = rel(uh, ref[:, :, k1s])
        ub = p.to_host(p.backward(p.to_device(ref[:, :, k1s].astype(o.complex))))
        e2 = rel(ub, u[:, x0s])
        # dealiased space
        upd = p.to_host(p.backward(p.to_device(ref[:, :, k1s].astype(o.complex)), padded=True))
        e3 = rel(upd, o._bwd_p(ref)[:, x0ps])
        # right-hand side and RK4 on a broadband field
        f0 = so.isotropic_field(o, seed=3, ncomp=nc)
        if solver == 'VV':
            f0 = o.cross2(o.K, f0)
        nu, eta, dt = 0.005, 0.01, 0.002
        if solver == 'NS':
            r_ref = o.ns_rhs(f0, nu, convection=conv or 'Vortex')
        elif solver == 'VV':
            r_ref = o.vv_rhs(f0, nu)
        else:
            r_ref = o.mhd_rhs(f0, nu, eta)
        d_u = p.to_device(f0[:, :, k1s])
        rhs = p.to_host(p.compute_rhs(p.empty_spectral(), d_u, nu, eta))
        e4 = rel(rhs, r_ref[:, :, k1s])
        u1, u2 = p.empty_spectral(), p.empty_spectral()
        for _ in range(2):
            p.rk4_step(d_u, u1, u2, dt, nu, eta)
        s_ref = o.solve(f0, solver, 2, dt, nu, eta=eta, convection=conv or 'Vortex')
        e5 = rel(p.to_host(d_u), s_ref[:, :, k1s])
        # energy: local parts sum to the global value
        t = torch.tensor([p.energy(d_u)], dtype=torch.float64, device='cuda')
        dist.all_reduce(t)
        e6 = abs(float(t.item()) - o.energy_fourier(s_ref))/o.energy_fourier(s_ref)
        errs = (e1, e2, e3, e4, e5, e6)
        ok = all(e < tol for e in errs) and not p.comm_timed_out()
        if rank == 0 or not ok:
            print('rank %d %s %s %s %s: %s %s' % (rank, N, prec, dealias, solver + ('/' + conv if conv else ''), ' '.join('%.1e' % e for e in errs),
                                                  'OK' if ok else 'FAIL'), flush=True)
        if not ok:
            fails.append((N, prec, dealias, solver, errs))
        del p
        dist.barrier()
    t = torch.tensor([len(fails)], device='cuda')
    dist.all_reduce(t)
    if rank == 0:
        print('SLAB_WORKER_RESULT fails=%d' % int(t.item()), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == '__main__':
    main()
