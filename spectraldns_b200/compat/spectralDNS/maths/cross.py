"""spectralDNS.maths.cross (reference maths/cross.py:16-35): cross1, cross2 under their submodule path."""
from . import cross1, cross2          # noqa: F401

__all__ = ['cross1', 'cross2']
