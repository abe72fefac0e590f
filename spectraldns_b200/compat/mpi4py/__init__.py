"""Drop-in for `from mpi4py import MPI` (solvers/spectralinit.py:11): one process per GPU; the
communicator is backed by torch.distributed when a process group is initialised."""
from . import MPI  # noqa
