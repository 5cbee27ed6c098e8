"""ctypes binding of lib/libhyperdb_b200.so (C ABI: include/hyperdb_b200.h).

The library is the product; if it is missing or cannot be loaded every compute call raises
(no Python/NumPy fallback exists in this package)."""
from __future__ import annotations

import ctypes as C
import os

HDB_F16, HDB_F32, HDB_F64 = 0, 1, 2
HDB_HOST, HDB_DEVICE = 0, 1
METRIC_IDS = {
    "dot_product": 0,
    "cosine_similarity": 1,
    "euclidean_metric": 2,
    "manhattan_distance": 3,
    "hamming_distance": 4,
    "jaccard_similarity": 5,
    "pearson_correlation": 6,
}
FLAG_FALLBACK, FLAG_QUERY_NAN, FLAG_TENSOR, FLAG_UNCERTIFIED, FLAG_EXCHANGE_ERROR = 1, 2, 4, 8, 16
SCORES_DISTANCE = 1

_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libhyperdb_b200.so")
_lib = None

i64, f64p, i64p, u32p, vp = C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.c_void_p

# name -> (restype, argtypes); must list every symbol of include/hyperdb_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "hdb_last_error": (C.c_char_p, []),
    "hdb_version": (C.c_int, []),
    "hdb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "hdb_matrix_create": (C.c_int, [C.c_int, C.c_int, i64, i64, i64, C.POINTER(vp)]),
    "hdb_matrix_destroy": (C.c_int, [vp]),
    "hdb_matrix_upload": (C.c_int, [vp, i64, i64, vp, C.c_int]),
    "hdb_matrix_adopt": (C.c_int, [vp, vp]),
    "hdb_matrix_finalize": (C.c_int, [vp]),
    "hdb_matrix_reserve": (C.c_int, [vp, i64]),
    "hdb_matrix_append": (C.c_int, [vp, i64, vp, C.c_int]),
    "hdb_matrix_remove_rows": (C.c_int, [vp, vp, i64, C.c_int]),
    "hdb_matrix_set_row_offset": (C.c_int, [vp, i64]),
    "hdb_matrix_set_stream": (C.c_int, [vp, vp]),
    "hdb_matrix_set_post_stream": (C.c_int, [vp, vp]),
    "hdb_matrix_set_sweep_overlap": (C.c_int, [vp, C.c_int]),
    "hdb_matrix_info": (C.c_int, [vp, C.POINTER(C.c_int), i64p, i64p, i64p, i64p]),
    "hdb_matrix_set_mask": (C.c_int, [vp, vp, C.c_int]),
    "hdb_matrix_set_row_order": (C.c_int, [vp, vp, C.c_int]),
    "hdb_matrix_set_range": (C.c_int, [vp, i64, i64]),
    "hdb_matrix_set_timestamps": (C.c_int, [vp, vp, C.c_int]),
    "hdb_matrix_kept_ts_max": (C.c_int, [vp, f64p, i64p]),
    "hdb_matrix_set_decay_reference": (C.c_int, [vp, C.c_double]),
    "hdb_matrix_stage1_recency": (C.c_int, [vp, C.c_double, C.c_double]),
    "hdb_query": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, i64, i64, C.c_double, vp, vp, vp, vp, C.c_int]),
    "hdb_scores": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]),
    "hdb_scores_ex": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "hdb_query_digest": (C.c_int, [vp, vp, C.c_int, C.c_int, i64, vp]),
    "hdb_query_digest_host": (C.c_int, [vp, C.c_int, i64, i64, vp]),
    "hdb_normalize_rows": (C.c_int, [C.c_int, C.c_int, i64, i64, vp, C.c_int, vp, C.c_int]),
    "hdb_merge_topk": (C.c_int, [C.c_int, vp, i64, i64, i64, i64, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int]),
    "hdb_exchange_create": (C.c_int, [C.c_int, C.c_int, C.c_int, i64, C.POINTER(vp)]),
    "hdb_exchange_destroy": (C.c_int, [vp]),
    "hdb_exchange_handle_bytes": (C.c_int, []),
    "hdb_exchange_local_handle": (C.c_int, [vp, vp]),
    "hdb_exchange_connect": (C.c_int, [vp, vp]),
    "hdb_exchange_connect_pointers": (C.c_int, [vp, C.POINTER(vp)]),
    "hdb_exchange_local_buffer": (C.c_int, [vp, C.POINTER(vp)]),
    "hdb_exchange_step": (C.c_int, [vp, vp, vp, i64, i64, i64, vp, vp, vp, vp]),
    "hdb_exchange_push": (C.c_int, [vp, vp, vp, i64]),
    "hdb_exchange_wait_merge": (C.c_int, [vp, vp, i64, i64, vp, vp, vp, vp]),
    "hdb_exchange_collect_async": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
    "hdb_exchange_stream": (C.c_int, [vp, C.POINTER(vp)]),
    "hdb_exchange_set_stream": (C.c_int, [vp, vp]),
    "hdb_matrix_attach_exchange": (C.c_int, [vp, vp]),
    "hdb_query_submit": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, i64, i64, C.c_double, C.c_int, i64p]),
    "hdb_query_collect": (C.c_int, [vp, i64, vp, vp, vp, vp]),
    "hdb_exchange_error": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "hdb_launch_count": (C.c_int64, [C.c_int]),
    "hdb_time_last_query": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "hdb_matrix_set_path": (C.c_int, [vp, C.c_int]),
    "hdb_matrix_set_max_group": (C.c_int, [vp, C.c_int]),
    "hdb_profile_enable": (C.c_int, [vp, C.c_int]),
    "hdb_profile_read": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
}


class NativeError(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


def lib():
    """Load the shared library once.  Raises NativeError (never falls back) if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeError(
                f"{_LIB_PATH} is missing: build it with `make -C local-hyperdb_b200` (or __graft_entry__.build()); "
                "hyperdb_b200 has no CPU fallback")
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    """Map a non-zero status to the exception the reference would raise."""
    if rc == 0:
        return
    msg = lib().hdb_last_error().decode()
    if "NaN" in msg or msg.startswith("Unknown metric"):
        raise ValueError(msg)
    raise NativeError(msg)
