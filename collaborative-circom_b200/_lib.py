"""ctypes binding of libcocg.so (the C ABI declared in include/cocg.h).

The library is the product; there is no CPU path.  Loading fails loudly when the shared object is missing
and `Context()` fails loudly when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcocg.so")

BN254, BLS12_381 = 0, 1
G1, G2 = 1, 2
OP_MUL, OP_ADD, OP_SUB, OP_NEG, OP_TO_MONT, OP_FROM_MONT = range(6)
EC_ADD, EC_MUL, EC_TO_AFFINE, EC_FROM_AFFINE, EC_NEG, EC_DBL = range(6)

# every symbol include/cocg.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "cocg_create", "cocg_destroy", "cocg_last_error", "cocg_version", "cocg_set_stream", "cocg_set_stream_priority", "cocg_sync",
    "cocg_launch_count", "cocg_malloc", "cocg_free", "cocg_h2d", "cocg_d2h", "cocg_memset0", "cocg_vec_op",
    "cocg_vec_scale_powers", "cocg_rep3_mul_local", "cocg_ntt", "cocg_bases_upload", "cocg_bases_free",
    "cocg_msm", "cocg_msm_host", "cocg_csr_upload", "cocg_csr_free", "cocg_spmv", "cocg_ec_op",
    "cocg_d2d", "cocg_host_alloc", "cocg_host_free", "cocg_rep3_mul_local_prf", "cocg_prf_fill", "cocg_prf_field_host",
    "cocg_bases_share", "cocg_csr_share", "cocg_bases_generate", "cocg_bases_download", "cocg_profile_enable",
    "cocg_profile_read", "cocg_profile_reset", "cocg_msm_multi", "cocg_vec_axpy", "cocg_csr_upload_form", "cocg_csr_download", "cocg_bases_generate_range",
    "cocg_fp_mul_ceiling", "cocg_vec_gather", "cocg_vec_scan", "cocg_vec_inv", "cocg_poly_eval", "cocg_vec_lincomb", "cocg_vec_fill",
    "cocg_msm_plan", "cocg_bases_check", "cocg_plonk_z_factors", "cocg_plonk_quotient_l1", "cocg_plonk_quotient_l2", "cocg_plonk_t_finish",
]

_lib = None


def msm_plan(curve: int, n: int):
    """(window bits, windows per scalar) of the table built for an n-point query."""
    c, w = ctypes.c_int(), ctypes.c_int()
    load().cocg_msm_plan(curve, n, ctypes.byref(c), ctypes.byref(w))
    return c.value, w.value


class CocgError(RuntimeError):
    pass


def load():
    """Load libcocg.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CocgError(f"{LIB_PATH} is missing: build it with `make -C {_HERE}/csrc` (there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, ci, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
    pvp = ctypes.POINTER(ctypes.c_void_p)
    sig = {
        "cocg_create": (ci, [pvp, ci, ci]),
        "cocg_destroy": (None, [vp]),
        "cocg_last_error": (ctypes.c_char_p, [vp]),
        "cocg_version": (ci, []),
        "cocg_set_stream": (ci, [vp, vp]),
        "cocg_set_stream_priority": (ci, [vp, ci]),
        "cocg_sync": (ci, [vp]),
        "cocg_launch_count": (u64, [vp]),
        "cocg_malloc": (ci, [vp, sz, pvp]),
        "cocg_free": (ci, [vp, vp]),
        "cocg_h2d": (ci, [vp, vp, vp, sz]),
        "cocg_d2h": (ci, [vp, vp, vp, sz]),
        "cocg_memset0": (ci, [vp, vp, sz]),
        "cocg_vec_op": (ci, [vp, ci, vp, vp, vp, sz]),
        "cocg_vec_scale_powers": (ci, [vp, vp, sz, vp, vp]),
        "cocg_rep3_mul_local": (ci, [vp, vp, vp, vp, vp, vp, vp, sz]),
        "cocg_ntt": (ci, [vp, pvp, ci, ctypes.c_uint, vp, ci, vp]),
        "cocg_bases_upload": (ci, [vp, ci, vp, sz, sz, ci, ctypes.POINTER(u64)]),
        "cocg_bases_free": (ci, [vp, u64]),
        "cocg_msm": (ci, [vp, u64, sz, sz, pvp, ci, ci, vp]),
        "cocg_msm_host": (ci, [vp, u64, sz, sz, pvp, ci, ci, vp]),
        "cocg_csr_upload": (ci, [vp, vp, vp, vp, sz, sz, ctypes.POINTER(u64)]),
        "cocg_csr_free": (ci, [vp, u64]),
        "cocg_spmv": (ci, [vp, u64, vp, sz, vp, vp]),
        "cocg_ec_op": (ci, [vp, ci, ci, vp, vp, vp]),
        "cocg_d2d": (ci, [vp, vp, vp, sz]),
        "cocg_host_alloc": (ci, [vp, sz, pvp]),
        "cocg_host_free": (ci, [vp, vp]),
        "cocg_rep3_mul_local_prf": (ci, [vp, vp, vp, vp, vp, vp, vp, ctypes.c_uint32, vp, sz]),
        "cocg_prf_fill": (ci, [vp, vp, ctypes.c_uint32, vp, sz]),
        "cocg_prf_field_host": (ci, [ci, vp, ctypes.c_uint32, u64, vp]),
        "cocg_bases_share": (ci, [vp, vp, u64, ctypes.POINTER(u64)]),
        "cocg_csr_share": (ci, [vp, vp, u64, ctypes.POINTER(u64)]),
        "cocg_bases_generate": (ci, [vp, ci, sz, vp, ctypes.POINTER(u64)]),
        "cocg_bases_generate_range": (ci, [vp, ci, sz, sz, vp, ctypes.POINTER(u64)]),
        "cocg_bases_download": (ci, [vp, u64, sz, sz, vp]),
        "cocg_profile_enable": (ci, [vp, ci]),
        "cocg_profile_read": (ci, [vp, ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)]),
        "cocg_profile_reset": (ci, [vp]),
        "cocg_fp_mul_ceiling": (ci, [vp, ci, ctypes.POINTER(ctypes.c_double)]),
        "cocg_msm_plan": (ci, [ci, sz, ctypes.POINTER(ci), ctypes.POINTER(ci)]),
        "cocg_bases_check": (ci, [vp, u64, ci, ctypes.POINTER(sz), ctypes.POINTER(sz)]),
        "cocg_vec_gather": (ci, [vp, vp, sz, vp, vp, sz]),
        "cocg_vec_scan": (ci, [vp, ci, vp, vp, sz]),
        "cocg_vec_inv": (ci, [vp, vp, vp, sz, ctypes.POINTER(sz)]),
        "cocg_poly_eval": (ci, [vp, vp, sz, vp, vp]),
        "cocg_vec_lincomb": (ci, [vp, ci, pvp, ctypes.POINTER(sz), vp, vp, sz]),
        "cocg_vec_fill": (ci, [vp, vp, sz, vp]),
        "cocg_plonk_z_factors": (ci, [vp, vp, sz, ci]),
        "cocg_plonk_quotient_l1": (ci, [vp, vp]),
        "cocg_plonk_quotient_l2": (ci, [vp, vp]),
        "cocg_plonk_t_finish": (ci, [vp, vp, vp, sz, vp, vp, vp]),
        "cocg_csr_upload_form": (ci, [vp, vp, vp, vp, sz, sz, ci, ctypes.POINTER(u64)]),
        "cocg_csr_download": (ci, [vp, u64, vp, vp, vp, ctypes.POINTER(sz)]),
        "cocg_vec_axpy": (ci, [vp, vp, vp, vp, vp, sz]),
        "cocg_msm_multi": (ci, [vp, ctypes.POINTER(u64), ctypes.POINTER(sz), ci, sz, pvp, ci, ci, pvp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
