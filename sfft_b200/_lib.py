"""
ctypes binding of libsfft_b200.so (include/sfft_b200.h).

The product path has no CPU fallback: if the CUDA library is missing or cannot be loaded this
module raises, loudly, at first use.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get('SFFTB_LIB') or os.path.join(HERE, 'libsfft_b200.so')     # SFFTB_LIB: A/B builds of the library

MEM_HOST, MEM_DEVICE = 0, 1
F64, F32 = 0, 1
STORE_F64, STORE_F32 = 0, 1

EXPORTS = ['sfftb_version', 'sfftb_last_error', 'sfftb_plan_create', 'sfftb_plan_destroy', 'sfftb_plan_dims',
           'sfftb_plan_set_stream', 'sfftb_plan_sync', 'sfftb_fit', 'sfftb_apply', 'sfftb_gss',
           'sfftb_export_normal_eq', 'sfftb_plan_set_timing', 'sfftb_timings', 'sfftb_last_solver',
           'sfftb_launch_count', 'sfftb_template_prepare', 'sfftb_template_state',
           'sfftb_template_mark_ready', 'sfftb_template_clone', 'sfftb_gss_template', 'sfftb_realize', 'sfftb_fits_decode', 'sfftb_fits_encode', 'sfftb_nan_union_fill', 'sfftb_nan_mask_apply',
           'sfftb_set_regularizer', 'sfftb_set_regularizer_varying', 'sfftb_gss_submit', 'sfftb_gss_template_submit', 'sfftb_gss_finish', 'sfftb_dbg_fft1d',
           'sfftb_dbg_row_spectra', 'sfftb_dbg_lag_tables', 'sfftb_plan_create_general', 'sfftb_export_solved_system',
           'sfftb_gss_submit_device', 'sfftb_gss_submit_delta', 'sfftb_gen_info', 'sfftb_plan_set_partition', 'sfftb_decorr', 'sfftb_convolve', 'sfftb_convolve_grid']


class Config(C.Structure):
    _fields_ = [('device', C.c_int), ('N0', C.c_int), ('N1', C.c_int), ('w0', C.c_int), ('w1', C.c_int),
                ('DK', C.c_int), ('DB', C.c_int), ('const_phot_ratio', C.c_int), ('storage', C.c_int),
                ('fold', C.c_int), ('sca_degree', C.c_int), ('reserved', C.c_int * 5)]


class Basis(C.Structure):
    """sfftb_basis: one tensor-product spatial basis as its 1-D tables (include/sfft_b200.h)."""
    _fields_ = [('nu', C.c_int), ('nv', C.c_int), ('nf', C.c_int), ('U', C.c_void_p), ('V', C.c_void_p),
                ('fu', C.c_void_p), ('fv', C.c_void_p)]


SCALING_ENTANGLED, SCALING_CONSTANT_DROP, SCALING_CONSTANT_SUM, SCALING_VARYING = 0, 1, 2, 3


class Dims(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('N0', 'N1', 'w0', 'w1', 'DK', 'DB', 'L0', 'L1', 'Fab', 'Fij', 'Fpq',
                                       'Fijab', 'NEQ', 'NEQ_FSfree', 'fold', 'sub_len')]


def build(verbose=False):
    """Compile libsfft_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ['make', '-j%d' % max(1, min(8, os.cpu_count() or 1)), '-C', os.path.join(HERE, 'csrc')]     # one translation unit per kernel family
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode:
        print(r.stdout)
    if r.returncode:
        raise RuntimeError('building libsfft_b200.so failed')
    return LIBPATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise ImportError('sfft_b200: %s is missing -- build it with `make -C sfft_b200/csrc` '
                          '(or python -c "import __graft_entry__ as g; g.build()"); there is no CPU fallback.' % LIBPATH)
    L = C.CDLL(LIBPATH)
    vp, ip, dp = C.c_void_p, C.c_int, C.POINTER(C.c_double)
    L.sfftb_version.restype = ip
    L.sfftb_last_error.restype = C.c_char_p
    L.sfftb_plan_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.sfftb_plan_destroy.argtypes = [vp]
    L.sfftb_plan_dims.argtypes = [vp, C.POINTER(Dims)]
    L.sfftb_plan_set_stream.argtypes = [vp, vp]
    L.sfftb_plan_sync.argtypes = [vp]
    L.sfftb_fit.argtypes = [vp, vp, vp, ip, ip, vp, ip]
    L.sfftb_apply.argtypes = [vp, vp, vp, ip, ip, vp, ip, vp, ip, ip]
    L.sfftb_gss.argtypes = [vp, vp, vp, vp, vp, ip, ip, vp, ip, vp, ip, ip]
    L.sfftb_export_normal_eq.argtypes = [vp, vp, vp]
    L.sfftb_template_prepare.argtypes = [vp, vp, vp, ip, ip]
    L.sfftb_template_state.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.sfftb_template_mark_ready.argtypes = [vp]
    L.sfftb_template_clone.argtypes = [vp, vp]
    L.sfftb_gss_template.argtypes = [vp, vp, vp, ip, ip, vp, ip, vp, ip, ip]
    L.sfftb_plan_set_timing.argtypes = [vp, ip]
    L.sfftb_timings.argtypes = [vp, C.POINTER(C.c_float), ip]
    L.sfftb_last_solver.argtypes = [vp]
    L.sfftb_launch_count.argtypes = [vp]
    L.sfftb_launch_count.restype = C.c_longlong
    L.sfftb_gss_submit.argtypes = [vp, vp, vp, vp, vp, ip, vp, vp, ip]
    L.sfftb_gss_template_submit.argtypes = [vp, vp, vp, ip, ip, vp, vp, ip]
    L.sfftb_gss_finish.argtypes = [vp]
    L.sfftb_fits_decode.argtypes = [ip, vp, vp, ip, ip, ip, C.c_double, C.c_double, vp, ip]
    L.sfftb_fits_encode.argtypes = [ip, vp, vp, ip, ip, ip, ip, vp]
    L.sfftb_nan_union_fill.argtypes = [ip, vp, vp, vp, vp, vp, ip, C.c_size_t, vp, vp]
    L.sfftb_nan_mask_apply.argtypes = [ip, vp, vp, ip, vp, C.c_size_t, C.c_double]
    L.sfftb_set_regularizer.argtypes = [vp, vp, vp, C.c_double]
    L.sfftb_set_regularizer_varying.argtypes = [vp, vp, vp]
    L.sfftb_realize.argtypes = [vp, vp, ip, vp, ip, ip, vp, vp, ip]
    L.sfftb_dbg_fft1d.argtypes = [ip, ip, ip, ip, vp, vp]
    L.sfftb_dbg_row_spectra.argtypes = [vp, ip, vp]
    L.sfftb_dbg_lag_tables.argtypes = [vp, vp, vp, vp, vp]
    L.sfftb_plan_create_general.argtypes = [C.POINTER(vp), C.POINTER(Config), C.POINTER(Basis), C.POINTER(Basis), C.POINTER(Basis), ip]
    L.sfftb_export_solved_system.argtypes = [vp, vp, vp]
    L.sfftb_gss_submit_device.argtypes = [vp, vp, vp, vp, vp, ip, vp, vp, ip]
    L.sfftb_gss_submit_delta.argtypes = [vp, vp, vp, C.c_longlong, vp, vp, C.c_longlong, vp, vp, ip, vp, vp, ip]
    L.sfftb_gen_info.argtypes = [vp, C.POINTER(C.c_int)]
    L.sfftb_plan_set_partition.argtypes = [vp, ip]
    L.sfftb_decorr.argtypes = [ip, vp, ip, ip, ip, vp, vp, vp, vp, C.c_double, ip, ip, ip, ip, vp, ip, vp]
    L.sfftb_convolve.argtypes = [ip, vp, vp, ip, ip, ip, vp, ip, ip, C.c_double, C.c_double, ip, ip, vp, ip]
    L.sfftb_convolve_grid.argtypes = [ip, vp, vp, ip, ip, ip, vp, ip, vp, ip, ip, C.c_double, ip, vp, ip]
    for name in EXPORTS:
        if name not in ('sfftb_last_error', 'sfftb_launch_count'):
            getattr(L, name).restype = ip
    _lib = L
    return L


class SFFTB200Error(Exception):
    def __init__(self, code, msg):
        super().__init__('MeLOn ERROR: %s' % msg)
        self.code = code


def check(rc):
    if rc != 0:
        msg = lib().sfftb_last_error().decode('utf-8', 'replace')
        if rc == -3:
            import numpy as np
            raise np.linalg.LinAlgError(msg)
        if rc == -4:
            raise ValueError(msg)
        raise SFFTB200Error(rc, msg)
