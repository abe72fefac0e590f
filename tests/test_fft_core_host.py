"""CPU test of the PRODUCT's FFT core (spectraldns_b200/csrc/fft_core.cuh): the same templates the CUDA kernels
instantiate -- radix plans, butterflies (radix 2/3/4/5/8/16), twiddle indexing, Stockham exchange maps, and the three
element types (float2 on the packed f32x2 intrinsics, double2, float2x2 = two columns per thread) -- compiled by g++
through tests/host/host_shim.h and driven with the threads of a line in lockstep (tests/host/fft_core_harness.cpp), for
every compiled transform length and every elements-per-thread choice, against a long-double DFT."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'tests', 'host', 'fft_core_harness.cpp')
BIN = os.path.join(ROOT, 'tests', 'host', 'build', 'fft_core_harness')
SIZES = [8, 12, 16, 32, 64, 128, 256, 512, 1024, 2048, 24, 48, 96, 192, 384, 768, 1536, 3072]


@pytest.fixture(scope='module')
def harness():
    deps = [SRC, os.path.join(ROOT, 'spectraldns_b200', 'csrc', 'fft_core.cuh'),
            os.path.join(ROOT, 'tests', 'host', 'host_shim.h')]
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(BIN), exist_ok=True)
        r = subprocess.run(['g++', '-std=c++20', '-O1', '-pthread', '-I', os.path.join(ROOT, 'tests', 'host'), '-o', BIN, SRC], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
    return BIN


@pytest.mark.parametrize('part,name,tol', [(0, 'float2', 6e-7), (1, 'double2', 1e-15), (2, 'float2x2', 6e-7)])
def test_fft_core_all_lengths_on_cpu(harness, part, name, tol):
    out = subprocess.run([harness, str(part)], stdout=subprocess.PIPE, text=True, timeout=600).stdout
    rows = [l.split() for l in out.strip().splitlines()]
    assert rows and all(r[0] == name for r in rows)
    seen = set()
    for typ, n, e, d, err in rows:
        assert float(err) < tol, (typ, n, e, d, err)
        seen.add(int(n))
    # every compiled length has at least one plan (sdns_size_supported / SDNS_SIZES in csrc/launch.cuh)
    assert set(SIZES) <= seen, sorted(set(SIZES) - seen)
    # forward and backward for every (N, E)
    assert len(rows) % 2 == 0 and len(rows) >= 2*len(SIZES)
