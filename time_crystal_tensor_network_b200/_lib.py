"""ctypes binding of the C ABI declared in include/tc_b200.h (libtc_b200.so, built in-tree).

There is no CPU fallback: if the library is missing, or a call fails, this raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TC_B200_LIB: another build of the same library (diagnostic / A-B builds); the default is the in-tree one
LIB_PATH = os.environ.get('TC_B200_LIB') or os.path.join(HERE, 'libtc_b200.so')

TRUNC_REFERENCE = 0
TRUNC_TEBD = 1
DBG_C, DBG_X, DBG_W, DBG_PERM = 0, 1, 2, 3
PROF_CLASSES = ('theta_gemm', 'qr', 'jacobi', 'finalize', 'bleft_gemm', 'measure', 'kick', 'reserved')


class EngineError(RuntimeError):
    pass


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int32)
_B = C.POINTER(C.c_int8)

# name -> (restype, argtypes); every symbol include/tc_b200.h declares
SIGNATURES = {
    'tc_version': (C.c_int, []),
    'tc_last_error': (C.c_char_p, []),
    'tc_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'tc_ctx_arena_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'tc_ctx_create': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, C.POINTER(_P)]),
    'tc_ctx_arena_bytes2': (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    'tc_ctx_create2': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, C.POINTER(_P)]),
    'tc_ctx_destroy': (C.c_int, [_P]),
    'tc_sync': (C.c_int, [_P]),
    'tc_ctx_info': (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'tc_get_flags': (C.c_int, [_P, _I]),
    'tc_set_product_state': (C.c_int, [_P, _B]),
    'tc_set_site': (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int, C.c_int]),
    'tc_get_site': (C.c_int, [_P, C.c_int, C.c_int, _D, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'tc_set_S': (C.c_int, [_P, C.c_int, C.c_int, _D, C.c_int]),
    'tc_get_S': (C.c_int, [_P, C.c_int, C.c_int, _D, C.POINTER(C.c_int)]),
    'tc_get_chi': (C.c_int, [_P, _I]),
    'tc_copy_chain': (C.c_int, [_P, C.c_int, _P, C.c_int]),
    'tc_get_trunc_err': (C.c_int, [_P, _D, C.c_int]),
    'tc_set_model': (C.c_int, [_P, _D, _D]),
    'tc_set_trunc': (C.c_int, [_P, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]),
    'tc_apply_layer': (C.c_int, [_P, C.c_int, C.c_int]),
    'tc_apply_kick': (C.c_int, [_P]),
    'tc_floquet_step': (C.c_int, [_P, C.c_int]),
    'tc_apply_two_site': (C.c_int, [_P, C.c_int, C.c_int, _D]),
    'tc_apply_one_site': (C.c_int, [_P, C.c_int, C.c_int, _D]),
    'tc_measure_dev': (C.c_int, [_P, _P, _P]),
    'tc_measure': (C.c_int, [_P, _D, _D]),
    'tc_overlap': (C.c_int, [_P, C.c_int, _P, C.c_int, _D]),
    'tc_correlation': (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _D, _D, _D]),
    'tc_floquet_run_dev': (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    'tc_floquet_run_host': (C.c_int, [_P, _D, _D, C.c_int, C.c_int, C.c_int, _D, _D, _D, _I]),
    'tc_profile': (C.c_int, [_P, C.c_int]),
    'tc_profile_read': (C.c_int, [_P, _D, C.POINTER(C.c_longlong), C.c_int]),
    'tc_dbg_get': (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t]),
    'tc_launch_count': (C.c_longlong, []),
    'tc_probe_fp64': (C.c_int, [C.c_int, C.c_int, _D]),
}

_lib = None


def load():
    """Load libtc_b200.so (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f'{LIB_PATH} is missing: build it with `python -m time_crystal_tensor_network_b200.build` '
            '(nvcc, sm_100a).  The engine has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=''):
    if status != 0:
        msg = load().tc_last_error()
        raise EngineError(f'{what}: {msg.decode() if msg else "unknown error"}')


def dptr(a):
    """double* view of a C-contiguous float64/complex128 numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(_D)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_I)
