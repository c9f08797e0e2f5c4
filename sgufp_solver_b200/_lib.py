"""ctypes binding of the C ABI (include/sgufp_b200.h).  Fails loudly when the CUDA library is
missing: there is no CPU fallback on the hot path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGUFP_B200_LIB") or os.path.join(_HERE, "libsgufp_b200.so")   # the override is for A/B timing of builds
_LIB = None

ERR = {0: "OK", -1: "ERR_ARG", -2: "ERR_MATCHING", -3: "ERR_CYCLIC", -4: "ERR_INSTANCE", -5: "ERR_LIMITS", -6: "ERR_CUDA"}

ip, i16p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_int16), C.POINTER(C.c_double)
u64p, i64p, u8p, vp = C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.c_void_p
cip = C.POINTER(C.c_int)

SIGNATURES = {
    "sgufp_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, ip, C.c_int, C.c_int, C.c_int64, C.c_int64]),
    "sgufp_destroy": (None, [vp]),
    "sgufp_last_error": (C.c_char_p, [vp]),
    "sgufp_dims": (C.c_int, [vp, cip, cip, cip]),
    "sgufp_network": (C.c_int, [vp, cip, cip, cip, i64p, i64p, ip, ip]),
    "sgufp_vbar_order": (C.c_int, [vp, ip]),
    "sgufp_processing_order": (C.c_int, [vp, ip]),
    "sgufp_slots": (C.c_int, [vp, ip, ip, ip, ip]),
    "sgufp_solve_path": (C.c_int, [vp, i16p, C.c_int, cip, dp, u64p, dp, cip, dp, dp, u8p, i64p]),
    "sgufp_solve_paths": (C.c_int, [vp, i16p, C.c_int, C.c_int, cip, dp, u64p, dp, cip, dp, dp, u8p, i64p]),
    "sgufp_partial_width": (C.c_int, [vp]),
    "sgufp_paths_partial": (C.c_int, [vp, i16p, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
    "sgufp_ray_partial": (C.c_int, [vp, i16p, C.c_int, C.c_int64, vp, vp]),
    "sgufp_finalize_paths": (C.c_int, [vp, i16p, C.c_int, C.c_int, i64p, i64p, cip, dp, u64p, dp, cip, dp]),
    "sgufp_create_sharded": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, ip, C.c_int, cip, C.c_int]),
    "sgufp_comm_unique_id": (C.c_int, [vp]),
    "sgufp_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "sgufp_comm_info": (C.c_int, [vp, cip, cip, cip, cip]),
    "sgufp_paths_reduced": (C.c_int, [vp, i16p, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp)]),
    "sgufp_stream": (vp, [vp]),
    "sgufp_cache_write": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, ip, C.c_int]),
    "sgufp_cache_dims": (C.c_int, [C.c_char_p, cip, cip, cip, cip, i64p]),
    "sgufp_create_from_cache": (C.c_int, [C.POINTER(vp), C.c_char_p, C.c_int, C.c_int64, C.c_int64]),
    "sgufp_cache_last_error": (C.c_char_p, []),
    "sgufp_clone": (C.c_int, [vp, C.POINTER(vp)]),
    "sgufp_cut_hash": (C.c_uint64, [u64p, dp, C.c_int]),
    "sgufp_last_stats": (C.c_int, [vp, cip, C.POINTER(C.c_float)]),
    "sgufp_run_length": (C.c_int, [vp, C.c_int]),
    "sgufp_last_kernel_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
}


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m sgufp_solver_b200.build` "
                "(or __graft_entry__.build()). There is no CPU fallback for the hot path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class SgufpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR.get(code, code)}: {msg}")
        self.code = code
