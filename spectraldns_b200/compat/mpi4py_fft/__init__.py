"""Import-time names only (utilities/__init__.py:13, demo/Isotropic.py:322-324)."""
from . import fftw  # noqa


def generate_xdmf(*args, **kwargs):
    """No h5py in the image: results are stored as .npz by spectraldns_b200.io; nothing to index."""
    return None
