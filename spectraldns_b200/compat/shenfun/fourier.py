"""shenfun.fourier.energy_fourier (tests/TG.py:5, demo/Isotropic.py:19) on the GPU."""
from spectraldns_b200.spaces import energy_fourier  # noqa
