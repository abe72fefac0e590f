"""Timer / MemoryUsage / create_profile with the reference's interface
(utilities/__init__.py:18-68, memoryprofiler.py:18-41, create_profile.py:9-85)."""
import os
import pstats
from time import time
from mpi4py import MPI

__all__ = ['Timer', 'MemoryUsage', 'create_profile', 'reset_profile']


class Timer(object):
    """Wall time of whole time steps: call once per step, final() reduces over ranks and prints
    Time / Fastest / Slowest exactly like the reference (its own step-time metric)."""

    def __init__(self):
        now = time()
        self.tic = self.t0 = now
        self.fastest_timestep = 1e8
        self.slowest_timestep = 0

    def __call__(self):
        now = time()
        step = now - self.t0
        self.t0 = now
        if step < self.fastest_timestep:
            self.fastest_timestep = step
        if step > self.slowest_timestep:
            self.slowest_timestep = step

    def final(self, verbose=True):
        comm = MPI.COMM_WORLD
        lo = tuple(comm.reduce(v, op=MPI.MIN, root=0) for v in (self.fastest_timestep, self.slowest_timestep))
        hi = tuple(comm.reduce(v, op=MPI.MAX, root=0) for v in (self.fastest_timestep, self.slowest_timestep))
        total = time() - self.tic
        if comm.Get_rank() == 0 and verbose:
            print('Time = {}'.format(total))
            print('Fastest = {}'.format(lo))
            print('Slowest = {}'.format(hi))


def _rss_vsz_mb():
    rss = vsz = 0
    try:
        with open('/proc/%d/status' % os.getpid()) as f:
            for line in f:
                if line.startswith('VmRSS:'):
                    rss = int(line.split()[1])//1024
                elif line.startswith('VmSize:'):
                    vsz = int(line.split()[1])//1024
    except OSError:
        pass
    return rss, vsz


class MemoryUsage(object):
    """MemoryUsage('label') prints resident / virtual memory summed over ranks."""

    def __init__(self, s):
        self.memory = 0
        self.memory_vm = 0
        self.first = True
        self(s)

    def __call__(self, s, verbose=True):
        prev, prev_vm = self.memory, self.memory_vm
        rss, vsz = _rss_vsz_mb()
        comm = MPI.COMM_WORLD
        self.memory = comm.reduce(rss) or 0
        self.memory_vm = comm.reduce(vsz) or 0
        if comm.Get_rank() == 0 and verbose:
            if self.first:
                print('Memory usage                    RSS accum     RSS total   Virtual  Virtual total')
                self.first = False
            print('{0:26s}  {1:10d} MB {2:10d} MB {3:10d} MB {4:10d} MB'.format(
                s, self.memory - prev, self.memory, self.memory_vm - prev_vm, self.memory_vm))


_PROFILED = ('integrate', 'rk4_step', 'ComputeRHS', 'compute_rhs', 'forward', 'backward', 'sync_to_host',
             'upload_state', 'update')


def create_profile(profiler):
    """Summarise a cProfile run: {name: (min_time, max_time, calls)} over ranks for the host
    functions of the time loop (the device work is asynchronous; use bench.py for kernel times)."""
    profiler.disable()
    stats = pstats.Stats(profiler).stats
    comm = MPI.COMM_WORLD
    out = {}
    for (fname, line, func), (cc, nc, tt, ct, callers) in stats.items():
        if func in _PROFILED:
            lo = comm.reduce(ct, op=MPI.MIN, root=0)
            hi = comm.reduce(ct, op=MPI.MAX, root=0)
            out[func] = (lo, hi, nc)
    if comm.Get_rank() == 0 and out:
        print('{0:20s} {1:>12s} {2:>12s} {3:>8s}'.format('function', 'min cumtime', 'max cumtime', 'calls'))
        for k, v in sorted(out.items()):
            print('{0:20s} {1:12.4e} {2:12.4e} {3:8d}'.format(k, v[0], v[1], v[2]))
    return out


def reset_profile(prof):
    prof.disable()
    prof.clear()
    prof.enable()


def inheritdocstrings(cls):
    """Class decorator: methods without a docstring take their parent's (utilities/__init__.py:71-80)."""
    import types
    for name, fn in vars(cls).items():
        if isinstance(fn, types.FunctionType) and not fn.__doc__:
            for base in cls.__mro__[1:]:
                doc = getattr(getattr(base, name, None), '__doc__', None)
                if doc:
                    fn.__doc__ = doc
                    break
    return cls


def cleanup():
    """Remove the result files of a run from the working directory (utilities/__init__.py:122-129): the reference's
    .h5 / .xdmf and this package's .npz archives."""
    import glob
    for f in glob.glob('*.h5') + glob.glob('*.xdmf') + glob.glob('*_c*.npz') + glob.glob('*_w*.npz'):
        try:
            os.remove(f)
        except OSError:
            pass
