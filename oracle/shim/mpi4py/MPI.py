"""TEST INFRASTRUCTURE ONLY -- single-rank stand-in for mpi4py.MPI."""
MIN, MAX, SUM = 'MIN', 'MAX', 'SUM'
C_FLOAT_COMPLEX, C_DOUBLE_COMPLEX, DOUBLE_COMPLEX, IN_PLACE = 'c8', 'c16', 'c16', None


class _Comm(object):
    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    def reduce(self, x, op=SUM, root=0):
        return x

    def allreduce(self, x, op=SUM):
        return x

    def bcast(self, x, root=0):
        return x

    def Reduce(self, a, b, op=SUM, root=0):
        b[...] = a

    def Allreduce(self, a, b, op=SUM):
        b[...] = a

    def Barrier(self):
        pass

    barrier = Barrier


COMM_WORLD = _Comm()
