"""HDF5File of the reference (h5io/HDF5File.py) on numpy archives; see spectraldns_b200/io.py."""
from spectraldns_b200.io import HDF5File  # noqa
__all__ = ['HDF5File']
