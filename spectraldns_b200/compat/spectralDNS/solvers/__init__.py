"""Solver modules of the B200 path: NS, VV, MHD (triply periodic)."""
