"""ctypes binding of the C ABI in include/b2r.h (libb2r.so, built by multimodal_rag_b200.build).

There is no CPU fallback: if the CUDA library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2R_LIB") or os.path.join(_HERE, "libb2r.so")     # B2R_LIB: development (A/B of two builds on one box)

B2R_OK, B2R_EINVAL, B2R_ECUDA, B2R_ENOMEM, B2R_EUNSUPPORTED = 0, 1, 2, 3, 4
SPACE_CODE = {"l2": 0, "cosine": 1, "ip": 2}
FLAG_NO_F32_MASTER = 1
TYPE_DEAD = 63

# every symbol include/b2r.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = (
    "b2r_abi_version", "b2r_last_error", "b2r_create", "b2r_destroy", "b2r_clear", "b2r_reserve",
    "b2r_ingest_f32", "b2r_tombstone", "b2r_query", "b2r_query_ex", "b2r_get_rows_f32", "b2r_count",
    "b2r_get_stats", "b2r_set_row_base", "b2r_merge_shards", "b2r_merge_shards_packed", "b2r_set_path", "b2r_launch_count",
    "b2r_set_kernel_timing", "b2r_kernel_time_ms", "b2r_save", "b2r_load", "b2r_column_set", "b2r_filter_eval", "b2r_query_async", "b2r_wait",
    "b2r_debug_trace",
    "b2r_idtab_create", "b2r_idtab_destroy", "b2r_idtab_clear", "b2r_idtab_live", "b2r_idtab_rows", "b2r_idtab_lookup",
    "b2r_idtab_append", "b2r_idtab_erase_rows", "b2r_idtab_ids_of",
    "b2r_xchg_create", "b2r_xchg_ipc_handle", "b2r_xchg_open", "b2r_xchg_push", "b2r_xchg_merge", "b2r_xchg_destroy", "b2r_query_push",
)


MAX_COLUMNS, WHERE_MAX_NODES = 16, 32
WHERE_LEAF, WHERE_AND, WHERE_OR = 0, 1, 2


class B2RWhereNode(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int32), ("column", ctypes.c_int32), ("lut_offset", ctypes.c_uint32),
                ("lut_values", ctypes.c_uint32)]


class B2RWhere(ctypes.Structure):
    _fields_ = [("n_nodes", ctypes.c_int32), ("nodes", B2RWhereNode * WHERE_MAX_NODES), ("lut", ctypes.c_void_p),
                ("lut_words", ctypes.c_int64)]


class B2RFilter(ctypes.Structure):
    _fields_ = [("type_mask", ctypes.c_uint64), ("allow_bits", ctypes.c_void_p), ("where", ctypes.POINTER(B2RWhere))]


class B2RStats(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int32), ("dim_padded", ctypes.c_int32), ("space", ctypes.c_int32),
                ("flags", ctypes.c_uint32), ("rows", ctypes.c_int64), ("live", ctypes.c_int64),
                ("capacity", ctypes.c_int64), ("bytes_device", ctypes.c_int64), ("n_queries", ctypes.c_int64),
                ("n_exact_fallbacks", ctypes.c_int64), ("sm_count", ctypes.c_int32), ("device", ctypes.c_int32),
                ("n_pool_queries", ctypes.c_int64), ("n_pool_entries", ctypes.c_int64)]


_lib = None


def load() -> ctypes.CDLL:
    """Load libb2r.so; raises RuntimeError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -m multimodal_rag_b200.build`); this engine has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32
    sigs = {
        "b2r_abi_version": (ctypes.c_int, []),
        "b2r_last_error": (ctypes.c_char_p, []),
        "b2r_create": (i32, [i32, i32, i64, i32, u32, ctypes.POINTER(vp)]),
        "b2r_destroy": (i32, [vp]),
        "b2r_clear": (i32, [vp]),
        "b2r_reserve": (i32, [vp, i64]),
        "b2r_ingest_f32": (i32, [vp, vp, i64, vp, ctypes.POINTER(i64), vp]),
        "b2r_tombstone": (i32, [vp, vp, i64, vp]),
        "b2r_query": (i32, [vp, vp, i32, i32, ctypes.POINTER(B2RFilter), vp, vp, vp, vp]),
        "b2r_query_ex": (i32, [vp, vp, i32, i32, ctypes.POINTER(B2RFilter), vp, vp, vp, vp, vp]),
        "b2r_get_rows_f32": (i32, [vp, vp, i64, vp, vp]),
        "b2r_count": (i64, [vp]),
        "b2r_get_stats": (i32, [vp, ctypes.POINTER(B2RStats)]),
        "b2r_set_row_base": (i32, [vp, i64]),
        "b2r_merge_shards": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp]),
        "b2r_merge_shards_packed": (i32, [vp, i64, i64, i64, i64, i32, i32, i32, vp, vp, vp, i32, vp]),
        "b2r_set_path": (i32, [vp, i32]),
        "b2r_launch_count": (i64, [vp]),
        "b2r_set_kernel_timing": (i32, [vp, i32]),
        "b2r_kernel_time_ms": (i32, [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64), i32]),
        "b2r_save": (i32, [vp, ctypes.c_char_p, vp]),
        "b2r_load": (i32, [ctypes.c_char_p, i32, i64, ctypes.POINTER(vp)]),
        "b2r_column_set": (i32, [vp, i32, i64, i64, vp, vp]),
        "b2r_filter_eval": (i32, [vp, ctypes.POINTER(B2RFilter), vp, vp]),
        "b2r_query_async": (i32, [vp, vp, i32, i32, ctypes.POINTER(B2RFilter), vp, vp, vp, vp, ctypes.POINTER(ctypes.c_uint64)]),
        "b2r_wait": (i32, [vp, ctypes.c_uint64]),
        "b2r_debug_trace": (i32, [vp, vp, i32, ctypes.POINTER(ctypes.c_int)]),
        "b2r_xchg_create": (i32, [i32, i32, i32, i32, i32, ctypes.POINTER(vp)]),
        "b2r_xchg_ipc_handle": (i32, [vp, vp]),
        "b2r_xchg_open": (i32, [vp, vp]),
        "b2r_xchg_push": (i32, [vp, vp, vp, vp, i32, i32, vp]),
        "b2r_xchg_merge": (i32, [vp, i32, i32, vp, vp, vp, vp]),
        "b2r_xchg_destroy": (i32, [vp]),
        "b2r_query_push": (i32, [vp, vp, vp, i32, i32, ctypes.POINTER(B2RFilter), vp, vp, vp, vp, vp, vp, vp]),
        "b2r_idtab_create": (i32, [i64, ctypes.POINTER(vp)]),
        "b2r_idtab_destroy": (i32, [vp]),
        "b2r_idtab_clear": (i32, [vp]),
        "b2r_idtab_live": (i64, [vp]),
        "b2r_idtab_rows": (i64, [vp]),
        "b2r_idtab_lookup": (i32, [vp, vp, vp, i64, i64, vp, ctypes.POINTER(i64)]),
        "b2r_idtab_append": (i32, [vp, vp, vp, i64, i64, i64, vp]),
        "b2r_idtab_erase_rows": (i32, [vp, vp, i64]),
        "b2r_idtab_ids_of": (i32, [vp, vp, i64, vp, i64, vp, ctypes.POINTER(i64)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "b2r") -> None:
    """Map a b2r_status to the exception the reference's callers expect
    (ValueError for bad arguments, RuntimeError for device failures)."""
    if rc == B2R_OK:
        return
    msg = (load().b2r_last_error() or b"").decode("utf-8", "replace")
    if rc == B2R_EINVAL:
        raise ValueError(msg or f"{what}: invalid argument")
    if rc == B2R_ENOMEM:
        raise MemoryError(msg or f"{what}: out of device memory")
    if rc == B2R_EUNSUPPORTED:
        raise NotImplementedError(msg or f"{what}: unsupported")
    raise RuntimeError(msg or f"{what}: CUDA failure (status {rc})")
