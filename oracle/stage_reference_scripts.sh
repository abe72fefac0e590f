#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Stage the reference's UNMODIFIED user-facing scripts (demos and test
# drivers) under baseline/_ref/ (git-ignored, travels to the GPU box) so that the -m gpu test
# tests/test_gpu_compat.py::test_reference_scripts_run_unchanged can run them on top of the B200
# implementation of `spectralDNS`.  Nothing is copied into tracked paths.
set -e
REF=${1:-/root/reference}
DST="$(dirname "$0")/../baseline/_ref"
mkdir -p "$DST/demo" "$DST/tests"
cp "$REF/demo/TG.py" "$REF/demo/TGMHD.py" "$REF/demo/Isotropic.py" "$DST/demo/"
cp "$REF/tests/TG.py" "$REF/tests/TGMHD.py" "$REF/tests/test_NSVV.py" "$REF/tests/test_MHD.py" "$DST/tests/"
echo "staged reference scripts in $DST"
