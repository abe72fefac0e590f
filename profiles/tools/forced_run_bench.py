"""Forced isotropic run through the drop-in layer, eager vs lazily mirrored state (1 GPU).

    python profiles/tools/forced_run_bench.py [--grid 256] [--precision double] [--dealias 2/3-rule] [--steps 10]

solve() with an update() callback made of the per-step expressions of demo/Isotropic.py:161-184 (numpy on the
context's arrays).  Eager: the state is copied to the host before every callback and back after it.
SDNS_LAZY_STATE=1: the same callback, answered on the device.  Prints ms per step (wall clock around solve(), which
synchronises at its end) and the number of state copies in each direction."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402


def run(a, lazy):
    os.environ['SDNS_LAZY_STATE'] = '1' if lazy else '0'
    from spectraldns_b200 import run as sdns_run
    sdns_run.activate()
    import spectralDNS
    from shenfun.fourier import energy_fourier
    config, get_solver, solve = spectralDNS.config, spectralDNS.get_solver, spectralDNS.solve
    dt = 0.0005
    config.update({'nu': 0.005428, 'dt': dt, 'T': dt*a.steps, 'convection': 'Vortex'})
    state = {}

    def update(c):
        if 'mask' not in state:
            state['mask'] = np.where(c.K2 <= 3**2, 1, 0)
        k2_mask = state['mask']
        c.U_hat[:, 0, 0, 0] = 0
        energy_new = energy_fourier(c.U_hat, c.T)
        energy_lower = energy_fourier(c.U_hat*k2_mask, c.T)
        alpha = np.sqrt((state['target'] - (energy_new - energy_lower))/energy_lower)
        c.U_hat *= (alpha*k2_mask + (1-k2_mask))
        state['e'] = energy_fourier(c.U_hat, c.T)

    m = str(int(round(np.log2(a.grid))))
    args = ['--M', m, m, m, '--precision', a.precision, '--dealias', a.dealias, 'NS']
    solver = get_solver(update=update, regression_test=lambda c: None, parse_args=args)
    c = solver.get_context()
    rng = np.random.RandomState(0)
    c.U[:] = 0.1*rng.standard_normal(c.U.shape)
    c.VT.forward(c.U, c.U_hat)
    state['target'] = energy_fourier(np.array(c.U_hat), c.T)
    config.params.t, config.params.tstep = 0.0, 0
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    solve(solver, c)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0)*1e3/a.steps
    dev = c['_dev']
    print('FORCED %s grid %d %s %s: %.2f ms/step  state copies h2d %d d2h %d  energy %.12e' % (
        'lazy ' if lazy else 'eager', a.grid, a.precision, a.dealias, ms, getattr(dev, 'h2d_copies', 0),
        getattr(dev, 'd2h_copies', 0), state['e']), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--grid', type=int, default=256)
    ap.add_argument('--precision', default='double')
    ap.add_argument('--dealias', default='2/3-rule')
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--lazy', type=int, default=0)
    run(ap.parse_args(), bool(ap.parse_args().lazy))
