"""ctypes binding of libpycmf_b200.so (C ABI declared in include/pycmf_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100a device is visible when a
context is created, this module raises instead of silently computing somewhere else.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpycmf_b200.so")

F32, F64 = 0, 1
LINEAR, LOGIT = 0, 1
ABI_VERSION = 1

_vp, _i32p, _i64, _int, _dbl = C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double

# name -> (restype, argtypes); mirrors include/pycmf_b200.h one to one
SIGNATURES = {
    "pycmf_abi_version": (_int, []),
    "pycmf_last_error": (C.c_char_p, []),
    "pycmf_create": (_int, [_int, C.POINTER(_vp)]),
    "pycmf_destroy": (_int, [_vp]),
    "pycmf_set_stream": (_int, [_vp, _vp]),
    "pycmf_set_option": (_int, [_vp, C.c_char_p, _dbl]),
    "pycmf_launch_count": (_i64, [_vp]),
    "pycmf_profile_enable": (_int, [_vp, _int]),
    "pycmf_profile_query": (_int, [_vp, C.c_char_p, C.POINTER(_dbl), C.POINTER(_i64)]),
    "pycmf_profile_reset": (_int, [_vp]),
    "pycmf_debug_tc_trace": (_int, [_vp, _vp, _i64]),
    "pycmf_gemm": (_int, [_vp, _int, _int, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _dbl, _dbl]),
    "pycmf_spmm": (_int, [_vp, _int, _i64, _i64, _i32p, _i32p, _vp, _vp, _i64, _i64, _vp, _i64, _dbl, _dbl]),
    "pycmf_resid_pass": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "pycmf_sqerr": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _int, _i32p, _i32p, _vp, _int, _vp]),
    "pycmf_mu_v_partial": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i32p, _i32p, _vp, _vp, _vp]),
    "pycmf_mu_v_apply": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _dbl, _dbl]),
    "pycmf_mu_left": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _int, _i32p, _i32p, _vp, _dbl, _dbl]),
    "pycmf_newton_left": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _int, _i32p, _i32p, _vp,
                                 _dbl, _dbl, _dbl, _int, _int, _dbl, _int, _i32p, _i64]),
    "pycmf_newton_v_xpart": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _i32p, _i32p, _vp,
                                    _int, _dbl, _i32p, _i64, _vp, _vp, C.POINTER(_int)]),
    "pycmf_newton_v_finish": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _int, _dbl, _dbl, _dbl,
                                     _i32p, _i64, _vp, _vp, _int, _int, _dbl]),
    "pycmf_safe_solve": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _dbl]),
    "pycmf_sample_indices": (_int, [_vp, _i64, _i64, _i64, C.c_uint64, C.c_uint64, _i32p]),
    "pycmf_topk_columns": (_int, [_vp, _int, _i64, _i64, _vp, _i64, _i64, _i32p]),
    "pycmf_sample_indices_sharded": (_int, [_vp, _i64, _i64, _i64, _i64, C.c_uint64, C.c_uint64, _i64, _i64, _i32p]),
}

_lib = None


def load():
    """dlopen the library and attach the prototypes. Raises ImportError when the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pycmf_b200: %s not found. Build it with `python -m pycmf_b200._build` (needs nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pycmf_abi_version() != ABI_VERSION:
        raise ImportError("pycmf_b200: ABI version mismatch between _lib.py and libpycmf_b200.so")
    _lib = lib
    return lib


def last_error():
    return load().pycmf_last_error().decode("utf-8", "replace")


class BackendError(RuntimeError):
    pass


def check(status):
    if status != 0:
        msg = last_error()
        if "Invalid link" in msg or "n_components" in msg or "hessian_pertubation" in msg:
            raise ValueError(msg)
        raise BackendError(msg)
