"""Names every solver module star-imports (reference solvers/spectralinit.py:8-62): communicator,
params, profiler, datatypes() and the default no-op hooks."""
import sys          # noqa: F401
import cProfile
import numpy as np
from mpi4py import MPI
from shenfun import CachedArrayDict as work_arrays   # noqa: F401
from spectralDNS import config
from spectralDNS.utilities import create_profile, MemoryUsage, Timer, reset_profile   # noqa: F401
from spectralDNS.h5io import HDF5File                 # noqa: F401
from spectralDNS.optimization import optimizer        # noqa: F401
from spectralDNS.maths import cross1, cross2, project, getintegrator   # noqa: F401

comm = MPI.COMM_WORLD
num_processes = comm.Get_size()
rank = comm.Get_rank()
params = config.params
profiler = cProfile.Profile()


def datatypes(precision):
    """(float, complex, mpitype) of a precision name."""
    table = {'single': (np.float32, np.complex64, MPI.C_FLOAT_COMPLEX),
             'double': (np.float64, np.complex128, MPI.C_DOUBLE_COMPLEX)}
    assert precision in table
    return table[precision]


def _default(fn):
    fn._sdns_default = True     # solve() skips the host refresh around untouched default hooks
    return fn


@_default
def regression_test(context):
    """Called once when solve() has finished."""


@_default
def update(context):
    """Called after every time step."""


@_default
def additional_callback(context):
    """Used by the adaptive integrators."""


def solve_linear(context):
    """Implicit solvers only."""


def conv(*args):
    raise NotImplementedError


def set_source(Source, **context):
    Source[:] = 0
    return Source


def end_of_tstep(context):
    """Return True to leave the time loop."""
    return False
