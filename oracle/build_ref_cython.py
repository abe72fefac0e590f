"""TEST INFRASTRUCTURE ONLY -- build the reference's own compiled kernels for this path into oracle/_ref/.

The reference's native layer for the hot path is three Cython templates
(/root/reference/spectralDNS/optimization/cython_{maths,solvers,integrators}.in) that its
optimization/setup.py expands once per precision (the `{0}` slot receives the three ctypedef lines
below) and cythonizes.  This recipe does the same expansion, reading the templates where they lie;
generated .pyx / .c files and the six extension modules go to oracle/_ref/ only (git-ignored, travels
to the GPU box).  Nothing is copied into tracked paths.

    python oracle/build_ref_cython.py            # needs /root/reference, Cython, numpy, g++

Users: tests/test_oracle.py::test_oracle_against_reference_cython_kernels pins the restatement in
oracle/sdns_oracle.py (cross1, cross2, add_pressure_diffusion, RK4 / ForwardEuler / AB2 stage algebra)
against these modules; nothing on the product path imports them.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
REF_OPT = '/root/reference/spectralDNS/optimization'

# optimization/setup.py:26-37 of the reference: the typedef block substituted for `{0}`
PRECISION = {
    'single': '\nctypedef np.complex64_t complex_t\nctypedef np.float32_t real_t\nctypedef np.int64_t int_t\n',
    'double': '\nctypedef np.complex128_t complex_t\nctypedef np.float64_t real_t\nctypedef np.int64_t int_t\n',
}
MODULES = ('maths', 'solvers', 'integrators')


def available():
    return all(os.path.exists(os.path.join(OUT, 'cython_%s_%s%s' % (p, m, sysconfig.get_config_var('EXT_SUFFIX'))))
               for p in PRECISION for m in MODULES)


def build(force=False):
    if not os.path.isdir(REF_OPT):
        return False                                   # the GPU box: use what was built in the container
    if available() and not force:
        return True
    import numpy as np
    from Cython.Build import cythonize                 # noqa: F401  (fail early if Cython is absent)
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    inc = ['-I' + sysconfig.get_paths()['include'], '-I' + np.get_include()]

    def one(job):
        prec, mod = job
        name = 'cython_%s_%s' % (prec, mod)
        pyx = os.path.join(OUT, name + '.pyx')
        with open(os.path.join(REF_OPT, 'cython_%s.in' % mod)) as f:
            src = f.read().format(PRECISION[prec])
        with open(pyx, 'w') as f:
            f.write(src)
        subprocess.run([sys.executable, '-m', 'cython', '-3', pyx, '-o', os.path.join(OUT, name + '.c')],
                       check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        subprocess.run(['gcc', '-O3', '-shared', '-fPIC', '-w', '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION'] + inc +
                       [os.path.join(OUT, name + '.c'), '-o', os.path.join(OUT, name + ext)],
                       check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        os.remove(os.path.join(OUT, name + '.c'))       # keep oracle/_ref small: it travels to the GPU box
        return name

    with ThreadPoolExecutor(max_workers=6) as ex:
        list(ex.map(one, [(p, m) for p in PRECISION for m in MODULES]))
    return available()


if __name__ == '__main__':
    ok = build(force='--force' in sys.argv)
    print('oracle/_ref: reference Cython kernels %s' % ('built' if ok else 'NOT available'))
