"""ctypes binding of libsdns_b200.so (include/sdns_b200.h).

The shared object is built in-tree by spectraldns_b200.build.  There is no fallback: if it is
missing, importing this module raises, and every entry point fails when no CUDA device exists.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get('SDNS_LIBPATH') or os.path.join(HERE, 'libsdns_b200.so')     # SDNS_LIBPATH: an experiment variant built with build.py --out

SDNS_ABI_VERSION = 1
SINGLE, DOUBLE = 0, 1
DEALIAS = {'None': 0, '2/3-rule': 1, '3/2-rule': 2}
SOLVER = {'NS': 0, 'VV': 1, 'MHD': 2}
CONVECTION = {'Vortex': 0, 'Divergence': 1, 'Standard': 2, 'Skewed': 3}
DECOMP = {'slab': 0, 'pencil': 1}
SPACE_T, SPACE_TP = 0, 1

# every symbol include/sdns_b200.h declares (tests check that the library exports them all)
SYMBOLS = ['sdns_abi_version', 'sdns_last_error', 'sdns_size_supported', 'sdns_plan_create',
           'sdns_plan_destroy', 'sdns_workspace_bytes', 'sdns_plan_set_workspace',
           'sdns_plan_set_stream', 'sdns_sync', 'sdns_local_shapes', 'sdns_k1_layout', 'sdns_comm_alloc', 'sdns_comm_handle',
           'sdns_comm_open', 'sdns_comm_status', 'sdns_forward', 'sdns_backward',
           'sdns_compute_rhs', 'sdns_compute_conv', 'sdns_rk4_step', 'sdns_euler_step', 'sdns_ab2_step', 'sdns_cross2', 'sdns_cross1', 'sdns_cross2_dense', 'sdns_project', 'sdns_add_pressure_diffusion', 'sdns_lincomb', 'sdns_errnorm',
           'sdns_energy', 'sdns_energy_weighted', 'sdns_scale_field', 'sdns_set_mode', 'sdns_enstrophy',
           'sdns_divergence_norm', 'sdns_spectrum', 'sdns_rk4_steps_host', 'sdns_launch_count',
           'sdns_profile_enable', 'sdns_profile_read', 'sdns_profile_read_nvlink', 'sdns_profile_read_copies', 'sdns_profile_timeline', 'sdns_xfer_stats']


class SdnsConfig(C.Structure):
    _fields_ = [('abi_version', C.c_int32),
                ('N', C.c_int32*3),
                ('L', C.c_double*3),
                ('precision', C.c_int32),
                ('dealias', C.c_int32),
                ('solver', C.c_int32),
                ('convection', C.c_int32),
                ('mask_nyquist', C.c_int32),
                ('decomposition', C.c_int32),
                ('kcut', C.c_int32*3),
                ('prune', C.c_int32),
                ('rank', C.c_int32),
                ('nranks', C.c_int32),
                ('device', C.c_int32),
                ('k1_layout', C.c_int32),
                ('reserved', C.c_int32*7)]


K1_LAYOUT = {'blocks': 0, 'cyclic': 1}


class Sdns2dConfig(C.Structure):
    _fields_ = [('abi_version', C.c_int32),
                ('N', C.c_int32*2),
                ('L', C.c_double*2),
                ('precision', C.c_int32),
                ('dealias', C.c_int32),
                ('solver', C.c_int32),
                ('mask_nyquist', C.c_int32),
                ('kcut', C.c_int32*2),
                ('device', C.c_int32),
                ('reserved', C.c_int32*8)]


SOLVER2D = {'NS2D': 0, 'Bq2D': 1}
SYMBOLS2D = ['sdns2d_last_error', 'sdns2d_plan_create', 'sdns2d_plan_destroy', 'sdns2d_workspace_bytes', 'sdns2d_plan_set_workspace',
             'sdns2d_plan_set_stream', 'sdns2d_sync', 'sdns2d_shapes', 'sdns2d_launch_count', 'sdns2d_forward', 'sdns2d_backward',
             'sdns2d_compute_rhs', 'sdns2d_rk4_step', 'sdns2d_euler_step', 'sdns2d_ab2_step', 'sdns2d_cross2',
             'sdns2d_add_pressure_diffusion']


def bind2d(L):
    """argtypes of the sdns2d_* entry points (doubly periodic solvers) on a loaded library."""
    vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
    L.sdns2d_last_error.restype = C.c_char_p
    L.sdns2d_plan_create.argtypes = [C.POINTER(vp), C.POINTER(Sdns2dConfig)]
    L.sdns2d_plan_destroy.argtypes = [vp]
    L.sdns2d_workspace_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.sdns2d_plan_set_workspace.argtypes = [vp, vp, C.c_size_t]
    L.sdns2d_plan_set_stream.argtypes = [vp, vp]
    L.sdns2d_sync.argtypes = [vp]
    L.sdns2d_shapes.argtypes = [vp, C.POINTER(C.c_int32*2), C.POINTER(C.c_int32*2), C.POINTER(C.c_int32*2)]
    L.sdns2d_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.sdns2d_forward.argtypes = [vp, i32, i32, vp, vp]
    L.sdns2d_backward.argtypes = [vp, i32, i32, vp, vp]
    L.sdns2d_compute_rhs.argtypes = [vp, vp, vp, dbl, dbl, dbl, vp, vp]
    L.sdns2d_rk4_step.argtypes = [vp, vp, vp, vp, dbl, dbl, dbl, dbl, vp]
    L.sdns2d_euler_step.argtypes = [vp, vp, vp, dbl, dbl, dbl, dbl, vp]
    L.sdns2d_ab2_step.argtypes = [vp, vp, vp, vp, dbl, i32, dbl, dbl, dbl, vp]
    L.sdns2d_cross2.argtypes = [vp, vp, vp]
    L.sdns2d_add_pressure_diffusion.argtypes = [vp, vp, vp, dbl, dbl, dbl, vp]
    for s in SYMBOLS2D:
        if s != 'sdns2d_last_error':
            getattr(L, s).restype = i32
    return L


class SdnsError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise SdnsError("%s is missing: run `python -m spectraldns_b200.build` (nvcc, sm_100a). "
                        "There is no CPU fallback." % LIBPATH)
    L = C.CDLL(LIBPATH)
    vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
    L.sdns_abi_version.restype = i32
    L.sdns_last_error.restype = C.c_char_p
    L.sdns_size_supported.argtypes = [i32, i32]
    L.sdns_plan_create.argtypes = [C.POINTER(vp), C.POINTER(SdnsConfig)]
    L.sdns_plan_destroy.argtypes = [vp]
    L.sdns_workspace_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.sdns_plan_set_workspace.argtypes = [vp, vp, C.c_size_t]
    L.sdns_plan_set_stream.argtypes = [vp, vp]
    L.sdns_sync.argtypes = [vp]
    L.sdns_comm_alloc.argtypes = [vp]
    L.sdns_comm_handle.argtypes = [vp, vp]
    L.sdns_comm_open.argtypes = [vp, vp, i32]
    L.sdns_comm_status.argtypes = [vp, C.POINTER(i32)]
    L.sdns_local_shapes.argtypes = [vp, C.POINTER(C.c_int32*3), C.POINTER(C.c_int32*3), C.POINTER(C.c_int32*3)]
    L.sdns_forward.argtypes = [vp, i32, i32, vp, vp]
    L.sdns_backward.argtypes = [vp, i32, i32, vp, vp]
    L.sdns_compute_rhs.argtypes = [vp, vp, vp, dbl, dbl, vp, vp]
    L.sdns_compute_conv.argtypes = [vp, vp, vp]
    L.sdns_rk4_step.argtypes = [vp, vp, vp, vp, dbl, dbl, dbl, vp]
    L.sdns_euler_step.argtypes = [vp, vp, vp, dbl, dbl, dbl, vp]
    L.sdns_ab2_step.argtypes = [vp, vp, vp, vp, dbl, i32, dbl, dbl, vp]
    L.sdns_cross2.argtypes = [vp, vp, vp, i32]
    L.sdns_cross1.argtypes = [vp, vp, vp, vp, C.c_longlong]
    L.sdns_cross2_dense.argtypes = [vp, vp, vp, vp]
    L.sdns_project.argtypes = [vp, vp]
    L.sdns_add_pressure_diffusion.argtypes = [vp, vp, vp, dbl, vp]
    L.sdns_lincomb.argtypes = [vp, vp, vp, i32, C.POINTER(dbl), C.POINTER(vp), i32]
    L.sdns_errnorm.argtypes = [vp, vp, vp, vp, dbl, dbl, i32, C.POINTER(dbl)]
    L.sdns_energy.argtypes = [vp, vp, i32, C.POINTER(dbl)]
    L.sdns_energy_weighted.argtypes = [vp, vp, i32, vp, i32, C.POINTER(dbl)]
    L.sdns_scale_field.argtypes = [vp, vp, i32, vp, i32, dbl, dbl]
    L.sdns_set_mode.argtypes = [vp, vp, i32, i32, i32, i32, dbl, dbl]
    L.sdns_enstrophy.argtypes = [vp, vp, C.POINTER(dbl)]
    L.sdns_divergence_norm.argtypes = [vp, vp, C.POINTER(dbl)]
    L.sdns_spectrum.argtypes = [vp, vp, i32, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.sdns_rk4_steps_host.argtypes = [vp, vp, vp, vp, vp, i32, dbl, dbl, dbl]
    L.sdns_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.sdns_k1_layout.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.sdns_profile_enable.argtypes = [vp, i32]
    L.sdns_profile_read.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(C.c_longlong), C.POINTER(dbl)]
    L.sdns_profile_read_nvlink.argtypes = [vp, i32, C.POINTER(dbl)]
    L.sdns_profile_read_copies.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(C.c_longlong)]
    L.sdns_profile_timeline.argtypes = [vp, C.POINTER(dbl), i32, C.POINTER(i32)]
    L.sdns_xfer_stats.argtypes = [vp, C.POINTER(dbl), C.POINTER(C.c_longlong)]
    for s in SYMBOLS:
        fn = getattr(L, s)
        if s not in ('sdns_last_error',):
            fn.restype = i32
    bind2d(L)
    if L.sdns_abi_version() != SDNS_ABI_VERSION:
        raise SdnsError('libsdns_b200.so ABI version mismatch')
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SdnsError('libsdns_b200: %s (status %d)' % (lib().sdns_last_error().decode(), rc))


def check2d(rc):
    if rc != 0:
        raise SdnsError('libsdns_b200: %s (status %d)' % (lib().sdns2d_last_error().decode(), rc))
