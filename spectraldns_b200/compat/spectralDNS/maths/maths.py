"""spectralDNS.maths.maths (reference maths/maths.py:8-11): project under its submodule path."""
from . import project                 # noqa: F401

__all__ = ['project']
