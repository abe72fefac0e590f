"""Run-time parameters of the triply periodic solvers: same option names, defaults, choices and
derived quantities as the reference's spectralDNS/config.py (v1.4.0), rebuilt from option tables.

What callers rely on (reference config.py line numbers):
  * config.params: a dict whose items are also attributes (:96-148); N = 2**M when no N was given
    (:110-118); dx = L/N (:111-112); nu, dt, Ri, Pr, eta come back cast to float32 in single
    precision (:123-127); M, N are stored as read-only int arrays and L accepts strings such as
    '2*pi' (:134-148)
  * config.triplyperiodic: argparse tree, solver name as positional sub-command (:212-239)
  * config.update(new, mesh): change defaults before parsing (:307-314); demos add their own
    arguments with config.triplyperiodic.add_argument (tests/TG.py:139-142)
  * config.doublyperiodic: the same for the 2-D solvers NS2D / Bq2D (:242-261)
The 'triplyperiodic' and 'doublyperiodic' meshes are on the B200 path; 'channel' raises.
"""
import argparse
import json
from collections import defaultdict

import numpy as np
from numpy import pi

__all__ = ['params', 'triplyperiodic', 'doublyperiodic', 'update', 'AttributeDict', 'Params', 'fft_plans']


class AttributeDict(dict):
    """dict with attribute access to its items: d.key is d['key'] (used for params and for the
    context returned by get_context())."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name)


_CAST_TO_PRECISION = frozenset(('nu', 'dt', 'Ri', 'Pr', 'eta'))
_INT_TRIPLES = frozenset(('M', 'N'))


def _eval_length(expr):
    return eval(str(expr), {'__builtins__': None}, {'pi': pi})


class Params(AttributeDict):
    """The global parameter dictionary."""

    def __getitem__(self, key):
        return dict.__getitem__(self, key)

    def __getattr__(self, name):
        if name in _CAST_TO_PRECISION and dict.__contains__(self, name):
            single = dict.get(self, 'precision', 'double') == 'single'
            return (np.float32 if single else np.float64)(dict.__getitem__(self, name))
        if dict.__contains__(self, name):
            return dict.__getitem__(self, name)
        if name == 'N':
            if not dict.__contains__(self, 'M'):
                raise KeyError('N')
            return 2**dict.__getitem__(self, 'M')
        if name == 'dx':
            return self.L/self.N
        raise KeyError(name)

    def __setitem__(self, key, value):
        if key in _INT_TRIPLES:
            value = np.array([int(str(v)) for v in value], dtype=int)
            value.flags.writeable = False
        elif key == 'L':
            value = np.array([_eval_length(v) for v in value], dtype=float)
            value.flags.writeable = False
        dict.__setitem__(self, key, value)

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v


fft_plans = defaultdict(lambda: 'FFTW_MEASURE', {'dct': 'FFTW_MEASURE'})


class _PlanAction(argparse.Action):
    """--planner_effort '{"fft": "FFTW_PATIENT"}': kept for command-line compatibility; the CUDA
    transforms have no planner."""
    def __call__(self, parser, namespace, values, option_string=None):
        fft_plans.update(json.loads(values))
        setattr(namespace, self.dest, fft_plans)


PlanAction = _PlanAction           # the reference's public name (config.py:153-158)


params = Params()
solver = None           # set by get_solver (reference __init__.py:63-64)
mesh = 'triplyperiodic'

# (flag, keyword arguments) -- reference config.py:164-209
_COMMON = [
    ('--precision', dict(default='double', choices=('single', 'double'))),
    ('--optimization', dict(default='', choices=('cython', 'weave', 'numba', 'pythran'),
                            help='Accepted and ignored: all kernels are sm_100a CUDA')),
    ('--make_profile', dict(default=0, type=int, help='Enable cProfile profiler')),
    ('--dt', dict(default=0.01, type=float, help='Time step size')),
    ('--T', dict(default=0.1, type=float, help='End time')),
    ('--write_result', dict(default=1e8, metavar='tstep', type=int, help='Write results every tstep')),
    ('--checkpoint', dict(default=1e8, type=int, help='Save intermediate result every...')),
    ('--nu', dict(default=0.000625, type=float, help='Viscosity')),
    ('--t', dict(default=0.0, type=float, help='Time')),
    ('--tstep', dict(default=0, type=int, help='Time step')),
    ('--filemode', dict(default='w', choices=('w', 'r', 'a'), help='Mode for opening result files')),
    ('--dealias', dict(default='2/3-rule', choices=('2/3-rule', '3/2-rule', 'None'),
                       help='Choose dealiasing method')),
    ('--decomposition', dict(default='slab', choices=('slab', 'pencil'),
                             help='Domain decomposition across GPUs')),
    ('--ntol', dict(default=7, type=int, help='Tolerance - number of accurate digits')),
    ('--threads', dict(default=1, type=int, help='Accepted and ignored (FFTW threads)')),
    ('--planner_effort', dict(action=_PlanAction, default=fft_plans, help='Accepted and ignored')),
    ('--h5filename', dict(default='results', type=str, help='Base name of checkpoint/result files')),
]
# reference config.py:212-226
_TRIPLY = [
    ('--convection', dict(default='Vortex', choices=('Standard', 'Divergence', 'Skewed', 'Vortex'),
                          help='Form of the nonlinear convective term')),
    ('--L', dict(default=[2*pi, 2*pi, 2*pi], metavar=('Lx', 'Ly', 'Lz'), nargs=3, help='Physical mesh size')),
    ('--M', dict(default=[6, 6, 6], metavar=('Mx', 'My', 'Mz'), nargs=3,
                 help='Mesh size is pow(2, M[i]) in direction i. Used if N is missing.')),
    ('--TOL', dict(type=float, default=1e-6, help='Tolerance for adaptive time integrator')),
    ('--integrator', dict(default='RK4', choices=('RK4', 'ForwardEuler', 'AB2', 'BS5_adaptive', 'BS5_fixed'),
                          help='Integrator for triply periodic domain')),
]
# sub-commands (reference config.py:228-239)
_SOLVERS = [
    ('NS', 'Regular Navier Stokes solver', []),
    ('VV', 'Velocity-Vorticity formulation', []),
    ('MHD', 'Magnetohydrodynamics solver', [('--eta', dict(default=0.01, type=float, help='MHD parameter'))]),
]


def _toggle(p, name, default, on_help, off_help):
    p.add_argument('--'+name, dest=name, action='store_true', help=on_help)
    p.add_argument('--no-'+name, dest=name, action='store_false', help=off_help)
    p.set_defaults(**{name: default})


# reference config.py:242-261
_DOUBLY = [
    ('--integrator', dict(default='RK4', choices=('RK4', 'ForwardEuler', 'AB2', 'BS5_fixed', 'BS5_adaptive'),
                          help='Integrator for doubly periodic domain')),
    ('--L', dict(default=[2*pi, 2*pi], nargs=2, metavar=('Lx', 'Ly'), help='Physical mesh size')),
    ('--convection', dict(default='Vortex', choices=('Vortex',), help='Form of the nonlinear convective term')),
    ('--TOL', dict(type=float, default=1e-6, help='Tolerance for adaptive time integrator')),
    ('--M', dict(default=[6, 6], nargs=2, metavar=('Mx', 'My'),
                 help='Mesh size is pow(2, M[i]) in direction i. Used if N is missing.')),
]
_SOLVERS_2D = [
    ('NS2D', 'Regular 2D Navier Stokes solver', []),
    ('Bq2D', 'Regular 2D Navier Stokes solver with Boussinesq model.',
     [('--Ri', dict(default=0.1, type=float, help='Richardson number')),
      ('--Pr', dict(default=1.0, type=float, help='Prandtl number'))]),
]


def _build():
    common = argparse.ArgumentParser(prog='spectralDNS', add_help=False)
    for flag, kw in _COMMON:
        common.add_argument(flag, **kw)
    _toggle(common, 'verbose', True, 'Print timings in the end', 'Do not print timings in the end')
    _toggle(common, 'mask_nyquist', True, 'Eliminate Nyquist frequency', 'Do not eliminate Nyquist frequency')
    meshes = []
    for options, solvers in ((_TRIPLY, _SOLVERS), (_DOUBLY, _SOLVERS_2D)):
        m = argparse.ArgumentParser(parents=[common])
        for flag, kw in options:
            m.add_argument(flag, **kw)
        sub = m.add_subparsers(dest='solver')
        for name, text, extra in solvers:
            sp = sub.add_parser(name, help=text)
            for flag, kw in extra:
                sp.add_argument(flag, **kw)
        meshes.append(m)
    return common, meshes[0], meshes[1]


parser, triplyperiodic, doublyperiodic = _build()


class _Unsupported(object):
    def __init__(self, name):
        self.name = name

    def __getattr__(self, attr):
        raise NotImplementedError("mesh '%s' is outside the B200 hot path (triply periodic NS/VV/MHD and doubly "
                                  "periodic NS2D/Bq2D only)" % self.name)


channel = _Unsupported('channel')


def update(new, mesh='triplyperiodic'):
    """Change parser defaults before parsing, e.g. config.update({'nu': 0.01, 'M': [5, 5, 5]})."""
    assert isinstance(new, dict)
    if 'planner_effort' in new:
        fft_plans.update(new['planner_effort'])
        new['planner_effort'] = fft_plans
    globals()[mesh].set_defaults(**new)
