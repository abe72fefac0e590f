"""Minimal h5py.File stand-in (h5py is not in the image) for the few calls demo/Isotropic.py makes
(:216-218, :286, :315-320): create_group / create_dataset / path indexing / close, kept in memory and
dumped to <name>.npz on close."""
import numpy as np


class _Group(dict):
    def __init__(self, root, path):
        dict.__init__(self)
        self._root, self._path = root, path

    def create_group(self, name):
        g = _Group(self._root, self._path + '/' + name)
        self[name] = g
        return g

    def create_dataset(self, name, data=None, **kw):
        self[name] = np.array(data)
        self._root._flat[(self._path + '/' + name).strip('/')] = self[name]
        return self[name]

    def __getitem__(self, key):
        node = self
        for part in key.strip('/').split('/'):
            node = dict.__getitem__(node, part)
        return node


class File(_Group):
    def __init__(self, name, mode='a', driver=None, comm=None, **kw):
        self._flat = {}
        _Group.__init__(self, self, '')
        self.filename = name
        self.mode = mode
        if mode in ('a', 'r', 'r+'):
            try:
                with np.load(name + '.npz', allow_pickle=False) as z:
                    for k in z.files:
                        node = self
                        parts = k.split('/')
                        for p in parts[:-1]:
                            node = node[p] if p in node else node.create_group(p)
                        node.create_dataset(parts[-1], data=z[k])
            except (IOError, OSError):
                pass

    def close(self):
        if self.mode != 'r' and self._flat:
            np.savez(self.filename + '.npz', **self._flat)
