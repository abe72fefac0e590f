"""TEST INFRASTRUCTURE ONLY -- build tests/host/build/libsdns_emu.so: the library's own sources
(spectraldns_b200/csrc/sdns_api.cu, inst.cu, passes.cuh, fft_core.cuh, launch.cuh) compiled by g++ against
tests/host/host_shim.h, an emulation of the CUDA execution model (one OS thread per CUDA thread, barriers,
warp-shuffle mailboxes, NaN-filled shared memory) and of the runtime calls the library makes.  Only a few small
transform lengths are instantiated.  Used by tests/test_kernels_emulated.py to check kernel logic without a GPU;
the product never loads it."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'spectraldns_b200', 'csrc')
OUT = os.path.join(HERE, 'build')
LIB = os.path.join(OUT, 'libsdns_emu.so')
SIZES = (8, 12, 16, 24, 32, 48, 64)
NFAM = 17


def build(extra=(), lib=LIB, sizes=SIZES, force=False):
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + [os.path.join(HERE, 'host_shim.h')]:
        with open(os.path.join(CSRC, f) if not os.path.isabs(f) else f, 'rb') as fh:
            h.update(fh.read())
    h.update(repr((tuple(extra), tuple(sizes))).encode())
    stamp = lib + '.stamp'
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return lib
    tag = os.path.basename(lib).replace('.so', '')
    flags = ['g++', '-std=c++20', '-O1', '-fPIC', '-pthread', '-w', '-DSDNS_HOST_SHIM', '-I', HERE, '-x', 'c++',
             '-DSDNS_SIZES(X)=' + ' '.join('X(%d)' % s for s in sizes if s % 5),
             '-DSDNS_SIZES_5(X)=' + ' '.join('X(%d)' % s for s in sizes if s % 5 == 0)] + list(extra)
    units = []
    for fam in range(NFAM):
        for prec in (32, 64):
            o = os.path.join(OUT, '%s_inst_%d_f%d.o' % (tag, fam, prec))
            units.append((o, flags + ['-DSDNS_FAMILY=%d' % fam, '-DSDNS_PREC=%d' % prec, '-c', os.path.join(CSRC, 'inst.cu'), '-o', o]))
    for api in ('sdns_api', 'sdns2d_api'):
        o = os.path.join(OUT, '%s_%s.o' % (tag, api))
        units.append((o, flags + ['-c', os.path.join(CSRC, api + '.cu'), '-o', o]))

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            raise RuntimeError(' '.join(cmd) + '\n' + r.stdout[-4000:])
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(lambda u: run(u[1]), units))
    run(['g++', '-shared', '-pthread', '-o', lib] + [u[0] for u in units])
    with open(stamp, 'w') as f:
        f.write(h.hexdigest())
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
