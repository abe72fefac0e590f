"""Build libsdns_b200.so in-tree with nvcc for sm_100a.

    python -m spectraldns_b200.build [--force] [--jobs N]

Every (kernel family, precision) pair of csrc/inst.cu is its own object file so the build
parallelises over the host cores; csrc/sdns_api.cu holds the plan and the C ABI.  The shared
object lands next to this file (git-ignored, but it travels to the GPU box with the tree).
"""
import os
import subprocess
import sys
import hashlib
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libsdns_b200.so')
NFAM = 17
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-std=c++17', '-O3', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
         '-Xfatbin=-compress-all']      # compressed cubins: the .so travels to the GPU box with the tree


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def _sources_digest(extra='', kernels_only=False):
    """Digest of the sources a unit depends on: the kernel instantiation units (inst.cu) do not include sdns_api.cu or
    the public header, so a change to the plan / C ABI recompiles one file instead of thirty."""
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + ['../../include/sdns_b200.h']:
        if kernels_only and f in ('sdns_api.cu', 'sdns2d_api.cu', '../../include/sdns_b200.h'):
            continue
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    h.update(extra.encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('command failed: %s\n%s' % (' '.join(cmd), r.stdout))
    return r.stdout


def build(force=False, jobs=None, verbose=False, sizes=None, out=None, precs=(32, 64), families=None):
    """Compile every CUDA translation unit for sm_100a and link libsdns_b200.so.
    families (experiment variants only): compile just these kernel families with SDNS_EXTRA_FLAGS and take every
    other object from the up-to-date default build."""
    global OBJ, LIB
    main_obj = OBJ
    if out:                                                      # experiment variant: its own objects and library
        OBJ = os.path.join(HERE, 'build', 'variant_' + os.path.basename(out))
        LIB = out
    os.makedirs(OBJ, exist_ok=True)
    extra = os.environ.get('SDNS_EXTRA_FLAGS', '').split()      # experiment switches, e.g. -DSDNS_NO_F0X
    if sizes:
        extra = extra + ['-DSDNS_SIZES(X)=' + ' '.join('X(%d)' % s for s in sizes if s % 5),
                         '-DSDNS_SIZES_5(X)=' + ' '.join('X(%d)' % s for s in sizes if s % 5 == 0)]
    digest = _sources_digest(' '.join(extra))
    stamp = os.path.join(OBJ, 'stamp')
    if (not force and os.path.exists(LIB) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest):
        return LIB
    nvcc = _nvcc()
    jobs = jobs or os.cpu_count() or 4
    units = []
    reuse = []
    kdigest = _sources_digest(' '.join(extra), kernels_only=True)
    kstamp = os.path.join(OBJ, 'stamp_kernels')
    kernels_fresh = (not force and os.path.exists(kstamp) and open(kstamp).read().strip() == kdigest)
    for fam in range(NFAM):
        for prec in (32, 64):
            if out and families is not None and fam not in families:
                reuse.append(os.path.join(main_obj, 'inst_%d_f%d.o' % (fam, prec)))
                continue
            o = os.path.join(OBJ, 'inst_%d_f%d.o' % (fam, prec))
            if kernels_fresh and os.path.exists(o):
                reuse.append(o)
                continue
            units.append((o, [nvcc] + ARCH + FLAGS + extra +
                          ['-DSDNS_FAMILY=%d' % fam, '-DSDNS_PREC=%d' % prec,
                           '-c', os.path.join(CSRC, 'inst.cu'), '-o', o]))
    for api in ('sdns_api', 'sdns2d_api'):          # plan + C ABI of the 3-D path, and of the 2-D solvers
        o_api = os.path.join(OBJ, api + '.o')
        if out and families is not None and 'api' not in families:
            reuse.append(os.path.join(main_obj, api + '.o'))
        else:
            units.append((o_api, [nvcc] + ARCH + FLAGS + extra + ['-c', os.path.join(CSRC, api + '.cu'), '-o', o_api]))
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        outs = list(ex.map(lambda u: _run(u[1]), units))
    if verbose:
        for o in outs:
            if o.strip():
                print(o)
    _run([nvcc] + ARCH + ['-shared', '-o', LIB] + [u[0] for u in units] + reuse)
    with open(stamp, 'w') as f:
        f.write(digest)
    with open(kstamp, 'w') as f:
        f.write(kdigest)
    return LIB


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--jobs', type=int, default=None)
    ap.add_argument('--verbose', action='store_true')
    ap.add_argument('--sizes', type=int, nargs='*', default=None)
    ap.add_argument('--out', default=None, help='write an experiment variant of the library here (with SDNS_EXTRA_FLAGS)')
    ap.add_argument('--families', type=int, nargs='*', default=None, help='variant builds: compile only these kernel families, reuse the default build for the rest')
    a = ap.parse_args()
    print(build(a.force, a.jobs, a.verbose, a.sizes, a.out, families=a.families))
