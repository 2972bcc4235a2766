"""ctypes binding of libcfmm_b200.so (the C ABI declared in include/cfmm_b200.h).

The library is the product; this module only loads it.  There is no Python or
CPU fallback: if the shared object is missing, or no CUDA device is visible when
a context is created, the failure is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CFMM_B200_LIB: development override (A/B runs of two builds of the same ABI on one box)
LIB_PATH = os.environ.get("CFMM_B200_LIB") or os.path.join(_HERE, "libcfmm_b200.so")

CFMM_OK = 0
CFMM_ERR_INVALID = -1
CFMM_ERR_CUDA = -2
CFMM_ERR_STATE = -3
CFMM_ERR_NOMEM = -4
CFMM_ERR_COMM = -5

POOL_PRODUCT = 0
POOL_GEOMEAN = 1
POOL_UNIV3 = 2

COMM_HANDLE_BYTES = 128

class SolveOpts(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("max_fun", C.c_int), ("pgtol", C.c_double), ("factr", C.c_double)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("fun_evals", C.c_int), ("status", C.c_int),
                ("f", C.c_double), ("pg_norm", C.c_double), ("solve_ms", C.c_double)]


# every symbol include/cfmm_b200.h declares: name -> (restype, argtypes)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_ctx = C.c_void_p
SYMBOLS = {
    "cfmm_create": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_int64]),
    "cfmm_destroy": (None, [_ctx]),
    "cfmm_last_error": (C.c_char_p, [_ctx]),
    "cfmm_version": (C.c_char_p, []),
    "cfmm_add_product": (C.c_int, [_ctx, C.c_int64, _dp, _dp, _ip]),
    "cfmm_add_geomean": (C.c_int, [_ctx, C.c_int64, _dp, _dp, _ip, _dp]),
    "cfmm_add_univ3": (C.c_int, [_ctx, C.c_int64, _dp, _dp, _ip, _ip, _dp, _dp]),
    "cfmm_pool_file_write": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, C.c_int64, _dp, _dp, _ip, _dp]),
    "cfmm_pool_file_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), _ip, _ip]),
    "cfmm_add_pool_file": (C.c_int, [_ctx, C.c_char_p]),
    "cfmm_finalize": (C.c_int, [_ctx]),
    "cfmm_num_pools": (C.c_int64, [_ctx]),
    "cfmm_num_tokens": (C.c_int64, [_ctx]),
    "cfmm_sweep": (C.c_int, [_ctx, _dp, _dp, _dp, C.c_int]),
    "cfmm_sweep_device": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cfmm_sweep_device_view": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cfmm_get_trades": (C.c_int, [_ctx, _dp, _dp]),
    "cfmm_solve": (C.c_int, [_ctx, _dp, _dp, _dp, _dp, C.POINTER(SolveOpts), _dp, C.POINTER(SolveInfo)]),
    "cfmm_update_reserves": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, _dp]),
    "cfmm_apply_trades": (C.c_int, [_ctx]),
    "cfmm_set_option": (C.c_int, [_ctx, C.c_char_p, C.c_int64]),
    "cfmm_last_sweep_ms": (C.c_int, [_ctx, C.POINTER(C.c_float)]),
    "cfmm_launch_count": (C.c_int64, [_ctx]),
    "cfmm_profile_read": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "cfmm_profile_read_times": (C.c_int, [_ctx, C.c_int, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int64)]),
    "cfmm_profile_reset": (C.c_int, [_ctx]),
    "cfmm_selftest_inrange_math": (C.c_int, [_ctx, _dp, _dp, C.c_int64, _ip]),
    "cfmm_debug_product_layout": (C.c_int, [C.c_int64, C.c_int64, _ip, C.c_int, C.c_int, C.c_int64, _ip,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_uint8), _ip]),
    "cfmm_debug_read_trace": (C.c_int, [_ctx, C.POINTER(C.c_uint64), C.c_int64, _ip]),
    "cfmm_host_alloc": (C.c_void_p, [C.c_size_t]),
    "cfmm_host_free": (None, [C.c_void_p]),
    "cfmm_comm_export": (C.c_int, [_ctx, C.c_void_p]),
    "cfmm_comm_attach": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_void_p]),
    "cfmm_comm_detach": (C.c_int, [_ctx]),
    "cfmm_comm_check": (C.c_int, [_ctx]),
}

_lib = None


class CFMMError(RuntimeError):
    """A non-zero status from libcfmm_b200 (the analogue of a Julia exception
    thrown by the reference: ArgumentError / BoundsError / MethodError)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libcfmm_b200 error {code}: {message}")
        self.code = code
        self.message = message


def load():
    """Load libcfmm_b200.so once and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  There is no CPU fallback for the sweep.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(lib, ctx, rc: int):
    if rc != CFMM_OK:
        msg = lib.cfmm_last_error(ctx)
        raise CFMMError(rc, msg.decode() if msg else "")
