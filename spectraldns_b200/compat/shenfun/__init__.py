"""Drop-in for the `shenfun` names the triply periodic solvers and demos import
(solvers/NS.py:8-9, MHD.py:8-9, spectralinit.py:12, h5io/HDF5File.py:4, demo/Isotropic.py:18),
implemented on the B200 plan.  See spectraldns_b200/spaces.py."""
from spectraldns_b200.spaces import (FunctionSpace, TensorProductSpace, VectorSpace, CompositeSpace,  # noqa
                                     Array, Function, CachedArrayDict)
from spectraldns_b200.io import ShenfunFile  # noqa
from . import fourier  # noqa
