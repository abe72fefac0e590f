"""TEST INFRASTRUCTURE ONLY -- single-rank stand-in for mpi4py (solvers/spectralinit.py:11,19-21)."""
from . import MPI  # noqa
