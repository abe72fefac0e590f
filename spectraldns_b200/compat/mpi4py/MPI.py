"""MPI names the hot path touches (spectralinit.py:19-21,28-29; utilities/__init__.py:58-62;
maths/integrators.py:92-98; tests/TG.py:121-122)."""
MIN, MAX, SUM = 'MIN', 'MAX', 'SUM'
C_FLOAT_COMPLEX, C_DOUBLE_COMPLEX, DOUBLE_COMPLEX, DOUBLE, FLOAT = 'c8', 'c16', 'c16', 'f8', 'f4'
IN_PLACE = None


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


class _Comm(object):
    def Get_size(self):
        d = _dist()
        return d.get_world_size() if d else 1

    def Get_rank(self):
        d = _dist()
        return d.get_rank() if d else 0

    def _all(self, x, op):
        d = _dist()
        if d is None:
            return x
        import numpy as np
        import torch
        a = np.asarray(x, dtype=np.float64)
        dev = 'cuda' if d.get_backend() == 'nccl' else 'cpu'
        t = torch.from_numpy(np.atleast_1d(a).copy()).to(dev)
        d.all_reduce(t, op={SUM: d.ReduceOp.SUM, MIN: d.ReduceOp.MIN, MAX: d.ReduceOp.MAX}[op])
        r = t.cpu().numpy()
        return r.reshape(a.shape) if a.ndim else type(x)(r[0]) if isinstance(x, (int, float)) else r[0]

    def allreduce(self, x, op=SUM):
        return self._all(x, op)

    def reduce(self, x, op=SUM, root=0):
        r = self._all(x, op)
        return r if self.Get_rank() == root else None

    def bcast(self, x, root=0):
        d = _dist()
        if d is None:
            return x
        box = [x]
        d.broadcast_object_list(box, src=root)
        return box[0]

    def Reduce(self, a, b, op=SUM, root=0):
        b[...] = self._all(a, op)

    def Allreduce(self, a, b, op=SUM):
        b[...] = self._all(a, op)

    def Barrier(self):
        d = _dist()
        if d is not None:
            d.barrier()

    barrier = Barrier


COMM_WORLD = _Comm()
