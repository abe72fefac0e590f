"""Per-pass device times of one configuration, for kernel experiments (1 GPU).

    python profiles/tools/passbench.py [--grid 256] [--precision double] [--dealias 2/3-rule] [--solver NS] [--reps 5]

Prints, for compute_conv (F0 writes the convection only), compute_rhs (F0 reads u_hat, writes rhs) and rk4_step
(F0 with the stage update), the CUDA-event time per launch and the algorithmic GB/s of every pass kernel.
SDNS_LIBPATH selects an experiment build of the library (spectraldns_b200/build.py --out).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
from spectraldns_b200.plan import Plan   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--grid', type=int, nargs='+', default=[256])
    ap.add_argument('--precision', default='double')
    ap.add_argument('--dealias', default='2/3-rule')
    ap.add_argument('--solver', default='NS')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--tag', default='')
    ap.add_argument('--configs', nargs='*', default=None,
                    help='several configurations in one process, each grid:precision:dealias[:solver], e.g. 512:double:2/3-rule')
    ap.add_argument('--only', default=None, help='conv | rhs | rk4: time just this call')
    a = ap.parse_args()
    if a.configs:
        for c in a.configs:
            f = c.split(':')
            a.grid, a.precision, a.dealias = [int(f[0])], f[1], f[2]
            a.solver = f[3] if len(f) > 3 else 'NS'
            run(a)
            torch.cuda.empty_cache()
    else:
        run(a)


def run(a):
    N = tuple(a.grid*3 if len(a.grid) == 1 else a.grid)
    p = Plan(N, precision=a.precision, dealias=a.dealias, solver=a.solver)
    g = torch.Generator(device='cuda').manual_seed(0)
    U = p.empty_physical()
    U.copy_(torch.randn(U.shape, generator=g, device='cuda', dtype=U.dtype)*0.1)
    u = p.forward(U)
    del U
    u1, u2, r = p.empty_spectral(), p.empty_spectral(), p.empty_spectral()
    out = {'grid': N, 'precision': a.precision, 'dealias': a.dealias, 'solver': a.solver, 'tag': a.tag,
           'lib': os.environ.get('SDNS_LIBPATH', 'default')}
    for name, fn in (('conv', lambda: p.compute_conv(r, u)), ('rhs', lambda: p.compute_rhs(r, u, 1e-3, 1e-3)),
                     ('rk4', lambda: p.rk4_step(u, u1, u2, 1e-4, 1e-3, 1e-3))):
        if a.only and name != a.only:
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        total = e0.elapsed_time(e1)/a.reps
        p.profile(True)
        for _ in range(a.reps):
            fn()
        prof = p.profile_read()
        p.profile(False)
        out[name] = {'ms': total, 'kernels': {k: {'us': v[0]/v[1]*1e3, 'GBps': v[2]/v[0]*1e-6, 'n': v[1]/a.reps}
                                                for k, v in prof.items()}}
        print('%-5s %s %s %s %s  %.3f ms  ' % (name, a.tag, N, a.precision, a.dealias, total) +
              '  '.join('%s %.0fus %.0fGB/s' % (k, v['us'], v['GBps']) for k, v in out[name]['kernels'].items()), flush=True)
    assert np.isfinite(p.energy(u))
    print('PASSBENCH ' + json.dumps(out))


if __name__ == '__main__':
    main()
