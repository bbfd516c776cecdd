"""ctypes binding of libspfresh_b200.so (include/spfresh_b200.h).

Thin by design: numpy arrays in, numpy arrays out, every call goes through the C ABI a Rust
`spann-cuda-sys` crate would bind.  There is no fallback: if the library is
not built or no B200 is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspfresh_b200.so")

SPF_OK = 0
METRIC_EUCLIDEAN, METRIC_MANHATTAN, METRIC_CHEBYSHEV = 0, 1, 2
ASSIGN_DEFAULT, ASSIGN_FORCE_EXACT, ASSIGN_NO_CSR = 0, 1, 2

# name -> (restype, argtypes); the single source for both binding and the export test
_u64p, _u32p, _f32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_float)
_vp, _vpp = C.c_void_p, C.POINTER(C.c_void_p)
SIGNATURES = {
    "spf_abi_version": (C.c_int, []),
    "spf_last_error": (C.c_char_p, []),
    "spf_device_count": (C.c_int, []),
    "spf_ctx_create": (C.c_int, [C.c_int, _vpp]),
    "spf_ctx_destroy": (None, [_vp]),
    "spf_ctx_device": (C.c_int, [_vp]),
    "spf_ctx_stream": (C.c_void_p, [_vp]),
    "spf_ctx_synchronize": (C.c_int, [_vp]),
    "spf_ctx_trim": (C.c_int, [_vp]),
    "spf_ctx_set_profiling": (C.c_int, [_vp, C.c_int]),
    "spf_ctx_kernel_ms": (C.c_float, [_vp, C.c_char_p]),
    "spf_ctx_launch_count": (C.c_uint64, [_vp]),
    "spf_ctx_last_overflow_rows": (C.c_uint32, [_vp]),
    "spf_ctx_set_param": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "spf_dataset_upload": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint64, _vpp]),
    "spf_dataset_from_device": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, _vpp]),
    "spf_dataset_free": (None, [_vp]),
    "spf_dataset_fetch_rows": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "spf_dataset_rows": (C.c_uint64, [_vp]),
    "spf_dataset_dim": (C.c_uint32, [_vp]),
    "spf_distance_pairs": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_uint32, C.c_uint64, _vp]),
    "spf_assign": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, _vp, C.c_uint32, C.c_float, C.c_int, _vpp]),
    "spf_assign_host": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, _vp, C.c_uint32, C.c_float,
                                  C.c_int, _vpp, _vpp]),
    "spf_assign_vectors": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, _vp, C.c_uint32, C.c_float, C.c_int, _vpp]),
    "spf_cluster_sums": (C.c_int, [_vp, _vp, _vp, _vp]),
    "spf_medoid_candidates": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "spf_assign_points": (C.c_uint64, [_vp]),
    "spf_assign_clusters": (C.c_uint32, [_vp]),
    "spf_assign_total": (C.c_uint64, [_vp]),
    "spf_assign_fetch": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "spf_assign_free": (None, [_vp]),
    "spf_update_medoids": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_uint32, _vp, _vp, _vp]),
    "spf_update_medoids_from": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "spf_kmpp_begin": (C.c_int, [_vp, C.c_int, C.c_uint64, _vpp]),
    "spf_kmpp_round": (C.c_int, [_vp, C.c_double, _u64p]),
    "spf_kmpp_push": (C.c_int, [_vp, C.c_uint64]),
    "spf_kmpp_rounds": (C.c_int, [_vp, C.POINTER(C.c_double), C.c_uint32, _u64p, C.POINTER(C.c_uint32)]),
    "spf_kmpp_last_sums": (C.c_int, [_vp, _f32p, C.POINTER(C.c_double)]),
    "spf_kmpp_begin_sharded": (C.c_int, [_vp, C.c_int, _vpp]),
    "spf_kmpp_fold_vector": (C.c_int, [_vp, _vp, _f32p]),
    "spf_kmpp_weight_total": (C.c_int, [_vp, C.c_float, C.POINTER(C.c_double)]),
    "spf_kmpp_pick_local": (C.c_int, [_vp, C.c_double, _u64p]),
    "spf_kmpp_set_vector": (C.c_int, [_vp, _vp]),
    "spf_kmpp_rounds_sharded": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(C.c_double), C.c_uint32, _u64p, C.POINTER(C.c_uint32)]),
    "spf_kmpp_free": (None, [_vp]),
    "spf_seq_sum_f32": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, _f32p]),
    "spf_farthest": (C.c_int, [_vp, C.c_int, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "spf_farthest_from": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, _vp, C.c_uint64, _f32p, _u64p]),
    "spf_index_pack": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vpp]),
    "spf_index_load_dir": (C.c_int, [_vp, C.c_char_p, _vp, C.c_uint32, C.c_uint32, _vpp]),
    "spf_index_save_dir": (C.c_int, [_vp, C.c_char_p]),
    "spf_index_free": (None, [_vp]),
    "spf_index_lists": (C.c_uint32, [_vp]),
    "spf_index_vectors": (C.c_uint64, [_vp]),
    "spf_search_batch": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, _vp, _vp, _vp,
                                   _vp, _vp]),
    "spf_search_sharded": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_float, _vp, _vp, _vp]),
    "spf_index_last_scan_bytes": (C.c_uint64, [_vp]),
    "spf_comm_unique_id": (C.c_int, [_vp]),
    "spf_comm_create": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vpp]),
    "spf_comm_destroy": (None, [_vp]),
    "spf_comm_world": (C.c_int, [_vp]),
    "spf_comm_rank": (C.c_int, [_vp]),
    "spf_kmeans_create": (C.c_int, [_vp, _vp, C.c_int, C.c_uint64, C.c_uint32, C.c_float, C.c_int, _vpp]),
    "spf_kmeans_set_centroids": (C.c_int, [_vp, _vp, _vp]),
    "spf_kmeans_step": (C.c_int, [_vp]),
    "spf_kmeans_fetch": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "spf_kmeans_assignment": (C.c_void_p, [_vp]),
    "spf_kmeans_set_balance": (C.c_int, [_vp, C.c_float]),
    "spf_assign_balanced": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, _vp, _vp, C.c_uint32, C.c_int, _vpp]),
    "spf_kmeans_free": (None, [_vp]),
    "spf_topk_merge": (C.c_int, [C.c_uint32, C.c_uint64, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class SpfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libspfresh_b200 error {code}: {msg}")
        self.code = code


def lib():
    """Loads the shared library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m spfresh_b200.build` "
                "(spfresh_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise SpfError(rc, lib().spf_last_error().decode("utf-8", "replace"))
    return rc


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def as_u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)
