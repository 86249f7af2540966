"""ctypes binding of libclann_b200.so — the C ABI declared in include/clann_b200.h.

This is the same stub a maintainer of the reference would write in Rust (`extern "C"` block);
Python stands in because this environment has no cargo. The library is REQUIRED: a missing or unloadable .so raises,
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_vp, _u32, _u64, _i32, _i64, _f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int64, C.c_float


class ClannConfig(C.Structure):
    """clann_config (include/clann_b200.h) <- Config (src/core/config.rs:17-35)."""
    _fields_ = [("num_tables", _u64), ("num_clusters_factor", _f32), ("k", _u64), ("delta", _f32)]


# clann_status
OK, ERR_DATA, ERR_CONFIG, ERR_CREATION, ERR_SEARCH, ERR_NOT_BUILT, ERR_BOUNDS, ERR_SERIALIZE, ERR_CUDA, ERR_ARG = (
    0, -1, -2, -3, -4, -5, -6, -7, -8, -9)

# clann_export_what
(X_NUM_CLUSTERS, X_CENTERS, X_ASSIGNMENT, X_RADII, X_OFFSETS, X_PERM, X_Q15, X_SKETCHES, X_TABLE_HASHES, X_TABLE_INDICES,
 X_BRUTE, X_NORMS, X_EST, X_QUERY_CODES, X_QUERY_SKETCHES, X_CLUSTER_ORDER, X_BUILD_MS, X_TABLE_DIR, X_REFERENCE_STREAM,
 X_QUERY_ANCHORS, X_QUERY_RANGES, X_STOP_POINTS, X_VISIT_LOG) = range(23)

# clann_allgather_fn / clann_allreduce_min_u64_fn (caller-supplied collectives of the cluster-sharded search)
ALLGATHER_FN = C.CFUNCTYPE(_i32, _vp, _vp, _vp, _u64, _vp)
ALLREDUCE_MIN_FN = C.CFUNCTYPE(_i32, _vp, _vp, _u64, _vp)

EXPORTED_SYMBOLS = [
    "clann_init_with_config", "clann_init_with_config_ex", "clann_set_option", "clann_set_delta", "clann_set_clustering", "clann_import_reference", "clann_set_functions",
    "clann_build", "clann_search", "clann_search_device", "clann_search_device_async", "clann_search_flush", "clann_search_async", "clann_search_wait", "clann_search_begin", "clann_search_step", "clann_state_bytes",
    "clann_state_ptr", "clann_search_merge", "clann_search_end", "clann_get_counters", "clann_export",
    "clann_comm_unique_id", "clann_comm_init", "clann_set_collectives", "clann_search_sharded", "clann_search_sharded_pair", "clann_search_sharded_multi", "clann_search_sharded_submit", "clann_search_sharded_flush", "clann_shard_stats",
    "clann_last_search_profile", "clann_tune", "clann_last_error", "clann_destroy", "clann_puffinn_search",
    "CPUFFINN_load_from_file", "CPUFFINN_index_create", "CPUFFINN_index_rebuild", "CPUFFINN_index_insert_cosine",
    "CPUFFINN_search_cosine", "CPUFFINN_get_distance_computations", "CPUFFINN_clear_distance_computations",
    "CPUFFINN_save_index",
]

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (never build implicitly on a GPU box: the .so ships with the repository snapshot)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: run `python -m clann_b200._build` (nvcc, sm_100a). clann_b200 has no CPU fallback.")
    L = C.CDLL(path)
    L.clann_last_error.restype = C.c_char_p
    L.clann_init_with_config.restype = _i32
    L.clann_init_with_config.argtypes = [_vp, _u64, _u32, C.POINTER(ClannConfig), C.POINTER(_vp)]
    L.clann_init_with_config_ex.restype = _i32
    L.clann_init_with_config_ex.argtypes = [_vp, _u64, _u32, C.POINTER(ClannConfig), _i32, _i32, C.POINTER(_vp)]
    L.clann_set_delta.restype, L.clann_set_delta.argtypes = _i32, [_vp, _f32]
    L.clann_set_option.restype, L.clann_set_option.argtypes = _i32, [_vp, C.c_char_p, _i64]
    L.clann_tune.restype, L.clann_tune.argtypes = _i32, [C.c_char_p, _i64]
    L.clann_set_clustering.restype, L.clann_set_clustering.argtypes = _i32, [_vp, _u64, _vp, _vp, _vp]
    L.clann_import_reference.restype, L.clann_import_reference.argtypes = _i32, [_vp, _u64, _vp, _u64]
    L.clann_set_functions.restype, L.clann_set_functions.argtypes = _i32, [_vp, _u64, _vp, _vp, _vp]
    L.clann_build.restype, L.clann_build.argtypes = _i32, [_vp]
    L.clann_search.restype, L.clann_search.argtypes = _i32, [_vp, _vp, _u64, _vp, _vp, _vp]
    L.clann_search_device.restype, L.clann_search_device.argtypes = _i32, [_vp, _vp, _u64, _vp, _vp, _vp, _vp]
    L.clann_search_device_async.restype, L.clann_search_device_async.argtypes = _i32, [_vp, _vp, _u64, _vp, _vp, _vp]
    L.clann_search_flush.restype, L.clann_search_flush.argtypes = _i32, [_vp, _vp]
    L.clann_search_async.restype, L.clann_search_async.argtypes = _i32, [_vp, _vp, _u64, _vp, _vp, _vp]
    L.clann_search_wait.restype, L.clann_search_wait.argtypes = _i32, [_vp]
    L.clann_search_begin.restype, L.clann_search_begin.argtypes = _i32, [_vp, _vp, _u64, _vp]
    L.clann_search_step.restype, L.clann_search_step.argtypes = _i32, [_vp, _vp]
    L.clann_state_bytes.restype, L.clann_state_bytes.argtypes = _u64, [_vp]
    L.clann_state_ptr.restype, L.clann_state_ptr.argtypes = _vp, [_vp]
    L.clann_search_merge.restype, L.clann_search_merge.argtypes = _i32, [_vp, _vp, _i32, C.POINTER(_u64), _vp]
    L.clann_search_end.restype, L.clann_search_end.argtypes = _i32, [_vp, _vp, _vp, _vp, _vp]
    L.clann_get_counters.restype, L.clann_get_counters.argtypes = _i32, [_vp, _u64, _vp, _vp, _vp]
    L.clann_comm_unique_id.restype, L.clann_comm_unique_id.argtypes = _i32, [_vp, _u64]
    L.clann_comm_init.restype, L.clann_comm_init.argtypes = _i32, [_vp, _i32, _i32, _vp]
    L.clann_set_collectives.restype, L.clann_set_collectives.argtypes = _i32, [_vp, ALLGATHER_FN, ALLREDUCE_MIN_FN, _vp]
    L.clann_search_sharded.restype, L.clann_search_sharded.argtypes = _i32, [_vp, _vp, _u64, _vp, _vp, _vp, _vp]
    L.clann_search_sharded_pair.restype = _i32
    L.clann_search_sharded_pair.argtypes = [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.clann_search_sharded_multi.restype = _i32
    L.clann_search_sharded_multi.argtypes = [_vp, _i32, _vp, _u64, _vp, _vp, _vp, _vp]
    L.clann_search_sharded_submit.restype = _i32
    L.clann_search_sharded_submit.argtypes = [_vp, _vp, _u64, _vp, _vp, _vp, _vp]
    L.clann_search_sharded_flush.restype = _i32
    L.clann_search_sharded_flush.argtypes = [_vp, _vp]
    L.clann_shard_stats.restype, L.clann_shard_stats.argtypes = _i32, [_vp, C.POINTER(_u64), C.POINTER(_u64), _vp]
    L.clann_export.restype, L.clann_export.argtypes = _i32, [_vp, _i32, _u64, _vp, _u64, C.POINTER(_u64)]
    L.clann_last_search_profile.restype, L.clann_last_search_profile.argtypes = _i32, [_vp, _vp, _vp]
    L.clann_destroy.restype, L.clann_destroy.argtypes = None, [_vp]
    # legacy ABI (libpuffinn-ffi/c_binder.h:14-26)
    L.CPUFFINN_load_from_file.restype, L.CPUFFINN_load_from_file.argtypes = _vp, [C.c_char_p, C.c_char_p]
    L.CPUFFINN_index_create.restype, L.CPUFFINN_index_create.argtypes = _vp, [C.c_char_p, _i32]
    L.CPUFFINN_index_rebuild.restype, L.CPUFFINN_index_rebuild.argtypes = _u64, [_vp, C.c_uint]
    L.CPUFFINN_index_insert_cosine.restype, L.CPUFFINN_index_insert_cosine.argtypes = None, [_vp, _vp, _i32]
    L.CPUFFINN_search_cosine.restype = C.POINTER(_u32)
    L.CPUFFINN_search_cosine.argtypes = [_vp, _vp, C.c_uint, _f32, _f32, _i32]
    L.CPUFFINN_get_distance_computations.restype, L.CPUFFINN_get_distance_computations.argtypes = C.c_uint, []
    L.CPUFFINN_clear_distance_computations.restype, L.CPUFFINN_clear_distance_computations.argtypes = None, []
    L.CPUFFINN_save_index.restype, L.CPUFFINN_save_index.argtypes = None, [_vp, C.c_char_p, _i32]
    L.clann_puffinn_search.restype = _i32
    L.clann_puffinn_search.argtypes = [_vp, _vp, _u32, _f32, _f32, _i32, _vp, _vp, _vp]
    _lib = L
    return L


def tune(key: str, value: int) -> None:
    """Process-global launch-parameter knob of the probe kernels (clann_tune); never changes a result."""
    if load().clann_tune(key.encode(), int(value)) != 0:
        raise RuntimeError(last_error())


def last_error() -> str:
    return load().clann_last_error().decode("utf-8", "replace")
