#!/bin/sh
# TEST / BASELINE INFRASTRUCTURE ONLY.  Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, travels to
# the GPU box; /root/reference does not exist there):
#   * its user-facing scripts (demos and test drivers), which tests/test_gpu_compat.py runs unchanged on top of the
#     B200 implementation of `spectralDNS`;
#   * its own pure-Python package for this path (spectralDNS/{__init__,config}.py, solvers/, maths/, utilities/,
#     h5io/, optimization/{__init__,cython_single,cython_double}.py) plus the reference's Cython kernels compiled by
#     oracle/build_ref_cython.py, dropped into optimization/ under the names its cython_{single,double}.py import:
#     this is what `bench.py --impl reference` times (over oracle/shim, the numpy stand-ins for the absent
#     shenfun / mpi4py / mpi4py-fft).
# `pip install --target baseline/_ref /root/reference` fails in this image (Cython 3 rejects shen/LUsolve.pyx, a
# channel-solver module that is out of scope), hence the explicit staging.  Nothing is copied into tracked paths.
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/../baseline/_ref"
mkdir -p "$DST/demo" "$DST/tests"
cp "$REF/demo/TG.py" "$REF/demo/TGMHD.py" "$REF/demo/Isotropic.py" "$REF/demo/TG2D.py" "$DST/demo/"
cp "$REF/tests/TG.py" "$REF/tests/TGMHD.py" "$REF/tests/test_NSVV.py" "$REF/tests/test_MHD.py" "$REF/tests/TG2D.py" \
   "$REF/tests/test_NS2D.py" "$DST/tests/"
PKG="$DST/spectralDNS"
rm -rf "$PKG"
mkdir -p "$PKG/optimization"
cp "$REF/spectralDNS/__init__.py" "$REF/spectralDNS/config.py" "$PKG/"
for d in solvers maths utilities h5io; do cp -r "$REF/spectralDNS/$d" "$PKG/$d"; done
cp "$REF/spectralDNS/optimization/__init__.py" "$REF/spectralDNS/optimization/cython_single.py" \
   "$REF/spectralDNS/optimization/cython_double.py" "$PKG/optimization/"
if ls "$HERE"/_ref/cython_*.so >/dev/null 2>&1; then cp "$HERE"/_ref/cython_*.so "$PKG/optimization/"; fi
find "$DST" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null || true
echo "staged the reference (scripts + package + compiled Cython kernels) in $DST"
