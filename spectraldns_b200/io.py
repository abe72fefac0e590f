"""Checkpoint / result files for the solve() loop.

The reference writes HDF5 through shenfun.ShenfunFile + h5py(mpio) (h5io/HDF5File.py:57-120); h5py
is not in this image, so the same two streams per run -- <name>_c (spectral checkpoint with `tstep`,
`t` attributes) and <name>_w (physical results) -- are stored as numpy .npz archives, one archive
per written time step and rank (nothing earlier than the step being written stays in memory):

    <name>_c[_rank<r>].npz                 latest checkpoint (overwritten), attrs tstep / t
    <name>_w[_rank<r>]_t<tstep>.npz        results of step tstep

Every archive carries the global shape and this rank's slice of each array (`meta__<key>`), so
`read_global` reassembles a field written on any number of ranks and `init_from_file` restarts on
a different rank count (the reference's init_from_file, h5io/HDF5File.py).  Cadence, the
kill-file protocol (`killspectraldns`, collective over all ranks) and the update_components hook
are the reference's.  This is side-band I/O, not part of the timed hot path.
"""
import glob
import os
import re
import sys
import numpy as np


class _Attrs(dict):
    def create(self, key, value):
        self[key] = value


class _Handle(object):
    def __init__(self, attrs):
        self.attrs = attrs


def _comm():
    from mpi4py import MPI          # the torch.distributed-backed stand-in (or the real one)
    return MPI.COMM_WORLD


class ShenfunFile(object):
    """ShenfunFile(name, space, mode=) with .open()/.close()/.f.attrs/.write(tstep, data, as_scalar=)."""
    def __init__(self, name, space=None, mode='w', per_step=False, **kw):
        comm = _comm()
        self.rank, self.nranks = comm.Get_rank(), comm.Get_size()
        self.stem = name + ('_rank%d' % self.rank if self.nranks > 1 else '')
        self.filename = self.stem + '.npz'
        self.per_step = per_step
        self.space = space
        self.mode = mode
        self._attrs = _Attrs()
        self.f = None
        if mode in ('r', 'a') and os.path.exists(self.filename):
            with np.load(self.filename, allow_pickle=False) as z:
                for k in z.files:
                    if k.startswith('attr__'):
                        self._attrs[k[6:]] = z[k].item()

    def open(self):
        self.f = _Handle(self._attrs)

    def close(self):
        self.f = None

    def _meta(self, arr):
        """(global shape, start, step of this rank's slice) of a local array of self.space (None: not distributed).  The step
        is 1 for the reference's contiguous slabs and the rank count on the axis-1 of a cyclic ownership (SDNS_K1_LAYOUT)."""
        sp = self.space
        try:
            spectral = np.iscomplexobj(arr)
            gshape = tuple(sp.global_shape(spectral)) if hasattr(sp, 'global_shape') else None
            sl = sp.local_slice(spectral)
            lead = arr.ndim - len(sl)
            start = (0,)*lead + tuple(int(s.start or 0) for s in sl)
            step = (1,)*lead + tuple(int(s.step or 1) for s in sl)
            if gshape is not None:
                gshape = tuple(arr.shape[:lead]) + gshape[-len(sl):]
            return gshape, start, step
        except Exception:
            return None, None, None

    def write(self, tstep, data, as_scalar=False):
        """data: {name: [array, (array, slices), ...]} as in h5io/HDF5File.py:25-52."""
        self.write_groups({tstep: data}, tstep)

    def write_groups(self, groups, tstep=0):
        """Several (tstep, data) groups in ONE archive, written once."""
        out = {}
        for ts, data in groups.items():
            self._collect(ts, data, out)
        for k, v in self._attrs.items():
            out['attr__' + k] = np.asarray(v)
        fn = ('%s_t%d.npz' % (self.stem, int(tstep))) if self.per_step else self.filename
        tmp = fn + '.tmp.npz'
        np.savez(tmp, **out)
        os.replace(tmp, fn)                      # a killed run never leaves a truncated checkpoint

    def _collect(self, tstep, data, out):
        for name, items in data.items():
            for item in items:
                if isinstance(item, tuple):
                    arr, sl = item
                    out['%s/slice/%s' % (name, tstep)] = np.array(np.asarray(arr)[tuple(sl)])
                else:
                    a = np.array(item)
                    key = '%s/3D/%s' % (name, tstep)
                    out[key] = a
                    gshape, start, step = self._meta(a)
                    if gshape is not None:
                        out['meta__' + key] = np.array(list(gshape) + list(start) + list(step), dtype=np.int64)


def read_global(name, key_prefix):
    """Reassemble the arrays whose key starts with key_prefix (e.g. 'U/3D/') from <name>.npz or from the per-rank
    files <name>_rank*.npz written by any number of ranks.  Returns ({key: global array}, attrs)."""
    files = sorted(glob.glob(name + '_rank*.npz'), key=lambda f: int(re.search(r'_rank(\d+)', f).group(1)))
    if not files and os.path.exists(name + '.npz'):
        files = [name + '.npz']
    if not files:
        raise IOError('no checkpoint %s[_rank*].npz' % name)
    out, attrs = {}, {}
    for fn in files:
        with np.load(fn, allow_pickle=False) as z:
            for k in z.files:
                if k.startswith('attr__'):
                    attrs[k[6:]] = z[k].item()
                elif k.startswith(key_prefix):
                    a = z[k]
                    if 'meta__' + k in z.files:
                        m = z['meta__' + k]
                        gshape, start = tuple(int(x) for x in m[:a.ndim]), tuple(int(x) for x in m[a.ndim:2*a.ndim])
                        step = tuple(int(x) for x in m[2*a.ndim:]) or (1,)*a.ndim        # older archives: contiguous blocks
                        if k not in out:
                            out[k] = np.zeros(gshape, dtype=a.dtype)
                        out[k][tuple(slice(s, s + (n - 1)*st + 1, st) for s, n, st in zip(start, a.shape, step))] = a
                    else:
                        out[k] = a
    return out, attrs


def init_from_file(name, u_hat, space):
    """Restart: fill this rank's block of the spectral state u_hat from the checkpoint <name>_c written by any
    number of ranks; returns (tstep, t)."""
    fields, attrs = read_global(name + '_c', '')
    keys = sorted(k for k in fields if '/3D/' in k)
    if not keys:
        raise IOError('checkpoint %s_c holds no 3D field' % name)
    g = fields[keys[-1]]
    sl = space.local_slice(True)
    u_hat[...] = g[(slice(None),)*(g.ndim - len(sl)) + tuple(sl)]
    return int(attrs.get('tstep', 0)), float(attrs.get('t', 0.0))


class HDF5File(object):
    """Per-step writer called from solve() (reference __init__.py:103; h5io/HDF5File.py:57-120)."""

    def __init__(self, filename, checkpoint={}, results={}):
        self.cfile = None
        self.wfile = None
        self.filename = filename
        self.checkpoint = checkpoint
        self.results = results
        self.before_host_read = None    # installed by the solver: brings the host mirrors up to date
        self.before_write = None        # installed by the solver: raises if the GPU exchange reported a fault

    def due(self, params):
        """True when this step may write something (the caller then syncs device -> host first).  Local and cheap:
        the collective decision is taken in update()."""
        return (params.tstep % params.write_result == 0 or params.tstep % params.checkpoint == 0
                or os.path.exists('killspectraldns'))

    def update(self, params, **kw):
        write = params.tstep % params.write_result == 0
        kill = self.check_if_kill()
        check = params.tstep % params.checkpoint == 0 or kill
        if not (write or check):
            return
        if self.before_write is not None:
            self.before_write()
        if self.before_host_read is not None:
            self.before_host_read()
        if self.cfile is None:
            self.cfile = ShenfunFile(self.filename + '_c', self.checkpoint.get('space'), mode=params.filemode)
            self.cfile.open()
            self.cfile.f.attrs.create('tstep', 0)
            self.cfile.f.attrs.create('t', 0.0)
            self.cfile.close()
        if self.wfile is None:
            self.wfile = ShenfunFile(self.filename + '_w', self.results.get('space'), mode=params.filemode, per_step=True)
        if write:
            self.update_components(**kw)
            self.wfile.write(params.tstep, self.results['data'], as_scalar=True)
        if check:
            self.cfile.open()
            self.cfile.f.attrs['tstep'] = int(params.tstep)
            self.cfile.f.attrs['t'] = float(params.t)
            self.cfile.close()
            # every checkpointed group ('0': current state, '1': previous one for multistep integrators) in one archive
            self.cfile.write_groups({int(key): val for key, val in self.checkpoint['data'].items()})
            if kill:
                sys.exit(1)

    def update_components(self, **kw):
        pass

    def open(self):
        if self.cfile:
            self.cfile.open()
        if self.wfile:
            self.wfile.open()

    def close(self):
        for f in (self.cfile, self.wfile):
            if f is not None and f.f:
                f.close()

    @staticmethod
    def check_if_kill(comm=None):
        """Collective over ALL ranks, like the reference (h5io/HDF5File.py: comm.allreduce(found)): a rank that has
        not seen the file yet must not keep stepping against peers that stop.  Rank 0 removes the file."""
        comm = comm if comm is not None else _comm()
        found = 1 if os.path.exists('killspectraldns') else 0
        if comm.Get_size() > 1:
            found = int(comm.allreduce(found))
        if found > 0:
            if comm.Get_rank() == 0:
                try:
                    os.remove('killspectraldns')
                except OSError:
                    pass
                print('killspectraldns Found! Stopping simulations cleanly by checkpointing...')
            if comm.Get_size() > 1:
                comm.Barrier()
            return True
        return False
