"""Checkpoint / result files for the solve() loop.

The reference writes HDF5 through shenfun.ShenfunFile + h5py(mpio) (h5io/HDF5File.py:57-120); h5py
is not in this image, so the same two files per run -- <name>_c (spectral checkpoint with `tstep`,
`t` attributes) and <name>_w (physical results) -- are stored as numpy .npz archives.  Cadence,
the kill-file protocol (`killspectraldns`) and the update_components hook are the reference's.
This is side-band I/O, not part of the timed hot path.
"""
import os
import sys
import numpy as np


class _Attrs(dict):
    def create(self, key, value):
        self[key] = value


class _Handle(object):
    def __init__(self, attrs):
        self.attrs = attrs


class ShenfunFile(object):
    """ShenfunFile(name, space, mode=) with .open()/.close()/.f.attrs/.write(tstep, data, as_scalar=)."""
    def __init__(self, name, space=None, mode='w', **kw):
        from .spaces import world
        rank, nranks, _ = world()
        self.filename = name + ('_rank%d' % rank if nranks > 1 else '') + '.npz'
        self.space = space
        self.mode = mode
        self._attrs = _Attrs()
        self._data = {}
        self.f = None
        if mode in ('r', 'a') and os.path.exists(self.filename):
            with np.load(self.filename, allow_pickle=False) as z:
                for k in z.files:
                    if k.startswith('attr__'):
                        self._attrs[k[6:]] = z[k].item()
                    else:
                        self._data[k] = z[k]

    def open(self):
        self.f = _Handle(self._attrs)

    def close(self):
        self.f = None

    def _flush(self):
        out = dict(self._data)
        for k, v in self._attrs.items():
            out['attr__' + k] = np.asarray(v)
        np.savez(self.filename, **out)

    def write(self, tstep, data, as_scalar=False):
        """data: {name: [array, (array, slices), ...]} as in h5io/HDF5File.py:25-52."""
        for name, items in data.items():
            for item in items:
                if isinstance(item, tuple):
                    arr, sl = item
                    key = '%s/slice/%s' % (name, tstep)
                    self._data[key] = np.array(np.asarray(arr)[tuple(sl)])
                else:
                    self._data['%s/3D/%s' % (name, tstep)] = np.array(item)
        self._flush()


class HDF5File(object):
    """Per-step writer called from solve() (reference __init__.py:103; h5io/HDF5File.py:57-120)."""

    def __init__(self, filename, checkpoint={}, results={}):
        self.cfile = None
        self.wfile = None
        self.filename = filename
        self.checkpoint = checkpoint
        self.results = results
        self.before_host_read = None    # installed by the solver: brings the host mirrors up to date

    def due(self, params):
        """True when this step writes something (the caller then syncs device -> host first)."""
        return (params.tstep % params.write_result == 0 or params.tstep % params.checkpoint == 0
                or 'killspectraldns' in os.listdir(os.getcwd()))

    def update(self, params, **kw):
        write = params.tstep % params.write_result == 0
        kill = self.check_if_kill(kw.get('comm_', None))
        check = params.tstep % params.checkpoint == 0 or kill
        if not (write or check):
            return
        if self.before_host_read is not None:
            self.before_host_read()
        if self.cfile is None:
            self.cfile = ShenfunFile(self.filename + '_c', self.checkpoint.get('space'), mode=params.filemode)
            self.cfile.open()
            self.cfile.f.attrs.create('tstep', 0)
            self.cfile.f.attrs.create('t', 0.0)
            self.cfile.close()
        if self.wfile is None:
            self.wfile = ShenfunFile(self.filename + '_w', self.results.get('space'), mode=params.filemode)
        if write:
            self.update_components(**kw)
            self.wfile.write(params.tstep, self.results['data'], as_scalar=True)
        if check:
            for key, val in self.checkpoint['data'].items():
                self.cfile.write(int(key), val)
                self.cfile.open()
                self.cfile.f.attrs['tstep'] = int(params.tstep)
                self.cfile.f.attrs['t'] = float(params.t)
                self.cfile.close()
                self.cfile._flush()
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
        found = 1 if 'killspectraldns' in os.listdir(os.getcwd()) else 0
        if comm is not None:
            found = comm.allreduce(found)
        if found > 0:
            if comm is None or comm.Get_rank() == 0:
                try:
                    os.remove('killspectraldns')
                except OSError:
                    pass
                print('killspectraldns Found! Stopping simulations cleanly by checkpointing...')
            return True
        return False
