"""Host-side mirror of the reference's public API for the build + search hot path.

Same names, argument meaning and error behaviour as the Rust crate (/root/reference/src/lib.rs:41-264):

    config = Config(num_tables=84, num_clusters_factor=0.4, k=10, delta=0.9, dataset_name="glove")
    index = init_with_config(AngularData(data), config)     # lib.rs:118   -> clann_init_with_config
    build(index)                                            # lib.rs:142   -> clann_build
    neighbours = search(index, query)                       # lib.rs:183   -> clann_search, [(distance, id), ...]

plus `search_batch`, the call a GPU wants (many queries per FFI crossing). Everything here is a thin shim over the
C ABI in include/clann_b200.h; no arithmetic of the hot path happens in Python.
"""
from __future__ import annotations

import ctypes as C
import os
import enum
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

__all__ = [
    "Config", "MetricsOutput", "AngularData", "ClusteredIndex", "ClusteredIndexError", "ConfigError", "DataError",
    "PuffinnCreationError", "PuffinnSearchError", "IndexNotFound", "IndexOutOfBounds", "SerializeError", "MetricsError",
    "CudaError", "init", "init_with_config", "build", "search", "search_batch", "PuffinnIndex", "get_recall_values",
    "brute_force_search", "generate_random_unit_vectors", "RunMetrics", "run_with_metrics", "serialize", "init_from_file",
]


# ------------------------------------------------------------------------------------------------ errors (core/errors.rs)

class ClusteredIndexError(Exception):
    """src/core/errors.rs:6-39"""


class ConfigError(ClusteredIndexError):
    def __str__(self):
        return f"Configuration Error: {self.args[0]}"


class DataError(ClusteredIndexError):
    def __str__(self):
        return f"Data Error: {self.args[0]}"


class PuffinnCreationError(ClusteredIndexError):
    def __str__(self):
        return f"PUFFINN Creation Error: {self.args[0]}"


class PuffinnSearchError(ClusteredIndexError):
    def __str__(self):
        return f"PUFFINN Search Error: {self.args[0]}"


class IndexNotFound(ClusteredIndexError):
    def __str__(self):
        return "Index Not Found Error"


class IndexOutOfBounds(ClusteredIndexError):
    def __str__(self):
        return f"Index Out of Bounds: {self.args[0]}"


class SerializeError(ClusteredIndexError):
    def __str__(self):
        return f"Serialize Error: {self.args[0]}"


class MetricsError(ClusteredIndexError):
    def __str__(self):
        return f"Metrics Error: {self.args[0]}"


class CudaError(ClusteredIndexError):
    """No reference counterpart: the reference has no GPU path. Raised when the device or the CUDA runtime fails."""


_STATUS_TO_ERROR = {
    _lib.ERR_DATA: DataError, _lib.ERR_CONFIG: ConfigError, _lib.ERR_CREATION: PuffinnCreationError,
    _lib.ERR_SEARCH: PuffinnSearchError, _lib.ERR_NOT_BUILT: IndexNotFound, _lib.ERR_BOUNDS: IndexOutOfBounds,
    _lib.ERR_SERIALIZE: SerializeError, _lib.ERR_CUDA: CudaError, _lib.ERR_ARG: ConfigError,
}


def _check(status: int) -> None:
    if status != _lib.OK:
        raise _STATUS_TO_ERROR.get(status, ClusteredIndexError)(_lib.last_error())


# ------------------------------------------------------------------------------------------------ Config (core/config.rs)

class MetricsOutput(enum.Enum):
    """src/core/config.rs:4-7. The sqlite sink itself is out of scope."""
    DB = "DB"
    NONE = "None"


@dataclass
class Config:
    """src/core/config.rs:17-35; defaults from :38-47."""
    num_tables: int = 10
    num_clusters_factor: float = 1.0
    k: int = 10
    delta: float = 0.9
    dataset_name: str = ""
    metrics_output: MetricsOutput = MetricsOutput.NONE

    @classmethod
    def new(cls, num_tables, num_clusters_factor, k, delta, dataset_name, metrics_output=MetricsOutput.NONE) -> "Config":
        """Config::new (config.rs:50-67)."""
        return cls(num_tables, num_clusters_factor, k, delta, dataset_name, metrics_output)

    def to_json_dict(self) -> dict:
        return {"num_tables": self.num_tables, "num_clusters_factor": self.num_clusters_factor, "k": self.k,
                "delta": self.delta, "dataset_name": self.dataset_name, "metrics_output": self.metrics_output.value}


# ------------------------------------------------------------------------------------------------ AngularData

class AngularData:
    """src/metricdata/angulardata.rs:6-62 — owns the n x d f32 matrix handed to the index. Norms and distances are
    computed on the device inside the index (k_row_norms, distance_point); this class only carries the rows."""

    def __init__(self, data):
        arr = np.ascontiguousarray(data, dtype=np.float32)
        if arr.ndim != 2:
            raise DataError("data must be a 2-D array")
        self.data = arr

    def num_points(self) -> int:
        return self.data.shape[0]

    def dimensions(self) -> int:
        return self.data.shape[1]

    def get_point(self, i: int) -> np.ndarray:
        return self.data[i]

    def subset(self, indices: Sequence[int]) -> "AngularData":
        return AngularData(self.data[np.asarray(indices, dtype=np.int64)])

    @staticmethod
    def similarity_type() -> str:
        return "angular"  # puffinn_types.rs:42-44


# ------------------------------------------------------------------------------------------------ ClusteredIndex

def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _Shape:
    """Stand-in for AngularData when the rows never lived in a host array of this process (device-resident / fp16 ingest)."""

    def __init__(self, n, d):
        self.n, self.d = n, d

    def num_points(self):
        return self.n

    def dimensions(self):
        return self.d


class ClusteredIndex:
    """ClusteredIndex<T> (src/core/index.rs:37-47), device resident behind the C ABI."""

    def __init__(self, config: Config, data: AngularData):
        if not isinstance(data, AngularData):
            data = AngularData(data)
        self.config = config
        self.data = data
        self._lib = _lib.load()
        self._h = C.c_void_p()
        cfg = _lib.ClannConfig(int(config.num_tables), float(np.float32(config.num_clusters_factor)), int(config.k),
                               float(np.float32(config.delta)))
        n, d = data.data.shape if data.data.size else (data.data.shape[0], data.data.shape[1] if data.data.ndim == 2 else 0)
        _check(self._lib.clann_init_with_config(_ptr(data.data) if n else None, n, d, C.byref(cfg), C.byref(self._h)))
        self.built = False

    @classmethod
    def from_rows(cls, config: Config, rows, n: int, d: int, dtype: str = "f32", on_device: bool = False) -> "ClusteredIndex":
        """clann_init_with_config_ex: rows as fp16 ("f16"; widened exactly on the device) and / or already on the device (`rows` is
        then a device pointer as an int; otherwise a numpy array of the matching dtype)."""
        self = cls.__new__(cls)
        self.config = config
        self.data = _Shape(n, d)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        cfg = _lib.ClannConfig(int(config.num_tables), float(np.float32(config.num_clusters_factor)), int(config.k),
                               float(np.float32(config.delta)))
        if on_device:
            ptr = C.c_void_p(int(rows))
        else:
            arr = np.ascontiguousarray(rows, np.float16 if dtype == "f16" else np.float32)
            if arr.shape != (n, d):
                raise DataError("rows do not have the stated shape")
            self._keep = arr
            ptr = _ptr(arr)
        _check(self._lib.clann_init_with_config_ex(ptr, n, d, C.byref(cfg), 1 if dtype == "f16" else 0, 1 if on_device else 0,
                                                   C.byref(self._h)))
        self.built = False
        return self

    # -- options / parity hooks (no counterpart in the reference API; used by tests and the multi-GPU driver)
    def set_option(self, key: str, value: int) -> None:
        _check(self._lib.clann_set_option(self._h, key.encode(), int(value)))

    def set_delta(self, delta: float) -> None:
        """Config.delta of a built index (stop rule only; no rebuild)."""
        _check(self._lib.clann_set_delta(self._h, float(np.float32(delta))))
        self.config.delta = float(delta)

    def set_clustering(self, centers, assignment, radii) -> None:
        c = np.ascontiguousarray(centers, np.uint64)
        a = np.ascontiguousarray(assignment, np.uint64)
        r = np.ascontiguousarray(radii, np.float32)
        _check(self._lib.clann_set_clustering(self._h, c.size, _ptr(c), _ptr(a), _ptr(r)))

    def import_reference(self, cluster: int, stream: bytes) -> None:
        buf = np.frombuffer(stream, np.uint8)
        _check(self._lib.clann_import_reference(self._h, cluster, _ptr(buf), buf.size))

    def set_functions(self, cluster: Optional[int], planes, signs, est) -> None:
        p = np.ascontiguousarray(planes, np.int16) if planes is not None else None
        s = np.ascontiguousarray(signs, np.int8) if signs is not None else None
        e = np.ascontiguousarray(est, np.float32) if est is not None else None
        cid = 0xFFFFFFFFFFFFFFFF if cluster is None else int(cluster)
        _check(self._lib.clann_set_functions(self._h, cid, _ptr(p), _ptr(s), _ptr(e)))

    # -- hot path
    def build(self) -> None:
        _check(self._lib.clann_build(self._h))
        self.built = True

    def search_batch(self, queries) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Returns (ids[nq,k] uint32 0xFFFFFFFF-padded, dists[nq,k] f32 +inf-padded, counts[nq])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.data.dimensions():
            raise PuffinnSearchError(f"query has {q.shape[1]} dimensions, index has {self.data.dimensions()}")
        nq, k = q.shape[0], int(self.config.k)
        ids = np.empty((nq, k), np.uint32)
        dists = np.empty((nq, k), np.float32)
        counts = np.empty(nq, np.uint32)
        _check(self._lib.clann_search(self._h, _ptr(q), nq, _ptr(ids), _ptr(dists), _ptr(counts)))
        return ids, dists, counts

    def search(self, query) -> List[Tuple[float, int]]:
        """ClusteredIndex::search (index.rs:311-439): [(distance, point id)] ascending, possibly fewer than k."""
        ids, dists, counts = self.search_batch(np.asarray(query, np.float32)[None, :])
        c = int(counts[0])
        return [(float(dists[0, i]), int(ids[0, i])) for i in range(c)]

    def counters(self, nq: int):
        cand, dc, vis = np.zeros(nq, np.uint64), np.zeros(nq, np.uint64), np.zeros(nq, np.uint32)
        _check(self._lib.clann_get_counters(self._h, nq, _ptr(cand), _ptr(dc), _ptr(vis)))
        return dict(candidates=cand, distance_computations=dc, clusters_visited=vis)

    def visit_log(self, nq: int) -> np.ndarray:
        """The per-visit rows of the last batch (option "visit_log" set before the search): uint32 [nq, cap, 4] =
        (cluster + 1, n_candidates, cluster_distance_computations, nanoseconds), zero rows beyond a query's visits."""
        log = self.export(_lib.X_VISIT_LOG, 0, np.uint32)
        return log.reshape(nq, -1, 4)

    def search_profile(self):
        ms = (C.c_float * 3)()
        launches = C.c_uint32(0)
        _check(self._lib.clann_last_search_profile(self._h, ms, C.byref(launches)))
        return dict(prep_ms=ms[0], probe_ms=ms[1], finish_ms=ms[2], launches=launches.value)

    def export(self, what: int, arg: int = 0, dtype=np.uint8) -> np.ndarray:
        size = C.c_uint64(0)
        _check(self._lib.clann_export(self._h, what, arg, None, 0, C.byref(size)))
        buf = np.empty(size.value, np.uint8)
        _check(self._lib.clann_export(self._h, what, arg, _ptr(buf), buf.size, C.byref(size)))
        return buf.view(dtype)

    @property
    def num_clusters(self) -> int:
        return int(self.export(_lib.X_NUM_CLUSTERS, 0, np.uint64)[0])

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.clann_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ lib.rs free functions

def init(data) -> ClusteredIndex:
    """lib.rs:76-82 — default Config."""
    return ClusteredIndex(Config(), data if isinstance(data, AngularData) else AngularData(data))


def init_with_config(data, config: Config) -> ClusteredIndex:
    """lib.rs:118-124"""
    return ClusteredIndex(config, data if isinstance(data, AngularData) else AngularData(data))


def build(index: ClusteredIndex) -> None:
    """lib.rs:142-148"""
    index.build()


def search(index: ClusteredIndex, query) -> List[Tuple[float, int]]:
    """lib.rs:183-189"""
    return index.search(query)


def search_batch(index: ClusteredIndex, queries):
    return index.search_batch(queries)


# ------------------------------------------------------------------------------------------------ persistence (index.rs:107-162,511-557)

_REC_MAGIC = b"CLB2REC\0"


def _write_record(f, name: str, payload: bytes) -> None:
    """One record of the flat container CPUFFINN_save_index writes (index.cu): magic, name length, reserved, payload length."""
    nb = name.encode()
    f.write(_REC_MAGIC + np.uint32(len(nb)).tobytes() + np.uint32(0).tobytes() + np.uint64(len(payload)).tobytes() + nb + payload)


def _read_records(path: str) -> dict:
    out = {}
    with open(path, "rb") as f:
        raw = f.read()
    pos = 0
    while pos < len(raw):
        if raw[pos:pos + 8] != _REC_MAGIC:
            raise SerializeError(f"{path} is not a libclann_b200 record file")
        name_len = int(np.frombuffer(raw[pos + 8:pos + 12], np.uint32)[0])
        payload_len = int(np.frombuffer(raw[pos + 16:pos + 24], np.uint64)[0])
        name = raw[pos + 24:pos + 24 + name_len].decode()
        start = pos + 24 + name_len
        if start + payload_len > len(raw):
            raise SerializeError(f"{path}: truncated record {name}")
        out[name] = raw[start:start + payload_len]   # the last record of a name wins (append-only file)
        pos = start + payload_len
    return out


def serialize(index: ClusteredIndex, directory_path: str) -> str:
    """lib.rs:255-264 -> ClusteredIndex::serialize (index.rs:511-557): the records the reference writes as HDF5 datasets —
    `config` (serde JSON of Config), `clusters` (serde JSON of Vec<ClusterCenter>: idx, center_idx, radius, assignment,
    brute_force, memory_used) and `index_{i}` = the bytes of puffinn::Index::serialize for every cluster that has a PUFFINN
    index — in the flat record container of CPUFFINN_save_index (there is no HDF5 here; only the container differs).
    Returns the file path (index_{dataset}_k{factor:.2}_L{tables}.clb2; the reference's name ends in .h5)."""
    import json
    if not os.path.isdir(directory_path):
        raise SerializeError(f"directory {directory_path} doesn't exist")   # index.rs:512-517
    if not index.built:
        raise SerializeError("index has not been built")
    cfg = index.config
    path = os.path.join(directory_path, "index_%s_k%.2f_L%d.clb2" % (cfg.dataset_name, cfg.num_clusters_factor, cfg.num_tables))
    K = index.num_clusters
    centers = index.export(_lib.X_CENTERS, 0, np.uint64)
    radii = index.export(_lib.X_RADII, 0, np.float32)
    offsets = index.export(_lib.X_OFFSETS, 0, np.uint64)
    perm = index.export(_lib.X_PERM, 0, np.uint32)
    brute = index.export(_lib.X_BRUTE, 0, np.uint8)
    clusters = [dict(idx=c, center_idx=int(centers[c]), radius=float(radii[c]),
                     assignment=perm[int(offsets[c]):int(offsets[c + 1])].tolist(), brute_force=bool(brute[c]), memory_used=0)
                for c in range(K)]
    config_json = json.dumps(dict(num_tables=int(cfg.num_tables), num_clusters_factor=float(cfg.num_clusters_factor), k=int(cfg.k),
                                  delta=float(cfg.delta), dataset_name=cfg.dataset_name, metrics_output="None"))
    with open(path, "wb") as f:
        _write_record(f, "config", config_json.encode())
        _write_record(f, "clusters", json.dumps(clusters).encode())
        for c in range(K):
            if not brute[c]:
                _write_record(f, f"index_{c}", index.export(_lib.X_REFERENCE_STREAM, c, np.uint8).tobytes())
    return path


def init_from_file(data, file_path: str) -> ClusteredIndex:
    """lib.rs:41-47 -> ClusteredIndex::new_from_file (index.rs:107-162): config and clusters from their JSON records, the hash
    functions of every cluster from its `index_{i}` stream; the device tables are rebuilt from them (0.3 s at the glove-100
    shape) instead of being copied — same functions, same rows, hence the same tables and the same answers."""
    import json
    if not os.path.exists(file_path):
        raise ConfigError(f"file {file_path} not found")   # index.rs:108-113
    rec = _read_records(file_path)
    try:
        c = json.loads(rec["config"].decode())
        clusters = json.loads(rec["clusters"].decode())
    except KeyError as e:
        raise ConfigError(f"record {e} missing in {file_path}")
    cfg = Config(int(c["num_tables"]), float(c["num_clusters_factor"]), int(c["k"]), float(c["delta"]), c.get("dataset_name", ""))
    index = init_with_config(data, cfg)
    n = index.data.num_points()
    assignment = np.zeros(n, np.uint64)
    for cl_ in clusters:
        assignment[np.asarray(cl_["assignment"], np.int64)] = cl_["idx"]
    index.set_clustering([cl_["center_idx"] for cl_ in clusters], assignment, [cl_["radius"] for cl_ in clusters])
    for cl_ in clusters:
        if not cl_["brute_force"]:
            name = f"index_{cl_['idx']}"
            if name not in rec:
                raise IndexNotFound()
            index.import_reference(cl_["idx"], rec[name])
    index.build()
    return index


# ------------------------------------------------------------------------------------------------ PuffinnIndex (legacy ABI)

class PuffinnIndex:
    """src/puffinn_binds/puffinn.rs:10-118 over the eight CPUFFINN_* symbols, exactly as the Rust crate drives them."""

    def __init__(self, raw, dim):
        self.raw, self.dim = raw, dim
        self.memory = 0

    @classmethod
    def new(cls, metric_data: AngularData, num_maps: int) -> Tuple["PuffinnIndex", int]:
        L = _lib.load()
        raw = L.CPUFFINN_index_create(metric_data.similarity_type().encode(), metric_data.dimensions())
        if not raw:
            raise PuffinnCreationError("Failed to create PUFFINN index")  # puffinn.rs:34-36
        self = cls(raw, metric_data.dimensions())
        for i in range(metric_data.num_points()):  # puffinn.rs:41-46
            row = np.ascontiguousarray(metric_data.get_point(i), np.float32)
            L.CPUFFINN_index_insert_cosine(raw, _ptr(row), self.dim)
        mem = L.CPUFFINN_index_rebuild(raw, num_maps)
        if mem == 0:
            raise PuffinnCreationError("Failed to create PUFFINN index, insufficient memory")  # puffinn.rs:52-54
        self.memory = mem
        return self, mem

    def search(self, query, k: int, max_dist: float, recall: float) -> List[int]:
        """puffinn.rs:77-118"""
        L = _lib.load()
        q = np.ascontiguousarray(query, np.float32)
        max_sim = float(np.float32(1.0) - np.float32(max_dist) / np.float32(2.0))  # puffinn_types.rs:77-79
        ptr = L.CPUFFINN_search_cosine(self.raw, _ptr(q), k, float(recall), max_sim, q.size)
        if not ptr:
            raise PuffinnSearchError("Search failed: returned null pointer.")
        try:
            if ptr[0] == 0xFFFFFFFF:
                return []
            return [int(ptr[i]) for i in range(k)]  # the crate reads exactly k words (puffinn.rs:108-114)
        finally:
            C.CDLL(None).free(ptr)

    FILTER_DEFAULT, FILTER_NONE, FILTER_SIMPLE = 0, 1, 2  # puffinn::FilterType (collection.hpp:22-34)

    def search_filtered(self, query, k: int, recall: float, filter_type: int = 0, max_sim: float = float("-inf")) -> Tuple[List[int], int]:
        """puffinn::Index::search(query, k, recall, max_sim, filter_type) (collection.hpp:324-334): the FilterType argument the
        reference's FFI never passes. Returns (ids best first, prefix length at which the stop rule fired or 0)."""
        L = _lib.load()
        q = np.ascontiguousarray(query, np.float32)
        if self.dim and q.size != self.dim:
            raise PuffinnSearchError(f"query has {q.size} dimensions, the index {self.dim}")
        ids = np.empty(max(k, 1), np.uint32)
        cnt, depth = C.c_uint32(0), C.c_uint32(0)
        st = L.clann_puffinn_search(self.raw, _ptr(q), k, float(recall), float(max_sim), int(filter_type), _ptr(ids), C.byref(cnt),
                                    C.byref(depth))
        if st != 0:
            raise PuffinnSearchError(_lib.last_error())
        return [int(v) for v in ids[: cnt.value]], int(depth.value)

    def save_to_file(self, file_path: str, index_id: int) -> None:
        """puffinn.rs:61-75 -> CPUFFINN_save_index: appends record "index_{id}" (the bytes of puffinn::Index::serialize,
        byte-compatible with the reference) to a flat record file; the reference uses an HDF5 dataset of the same name."""
        _lib.load().CPUFFINN_save_index(self.raw, os.fspath(file_path).encode(), int(index_id))

    @classmethod
    def new_from_file(cls, file_path: str, index_id: int) -> "PuffinnIndex":
        """puffinn.rs:121-141 -> CPUFFINN_load_from_file: the stored index answers queries without being rebuilt."""
        raw = _lib.load().CPUFFINN_load_from_file(os.fspath(file_path).encode(), f"index_{int(index_id)}".encode())
        if not raw:
            raise SerializeError(f"could not load index_{index_id} from {file_path}: {_lib.last_error()}")
        return cls(raw, 0)


def get_distance_computations() -> int:
    return int(_lib.load().CPUFFINN_get_distance_computations())


def clear_distance_computations() -> None:
    _lib.load().CPUFFINN_clear_distance_computations()


# ------------------------------------------------------------------------------------------------ measurement helpers (utils/mod.rs)

def get_recall_values(dataset_distances: np.ndarray, run_distances: Sequence[Sequence[float]], count: int):
    """src/utils/mod.rs:59-95 — a returned distance counts if it is <= the count-th true distance + 1e-3."""
    recalls = []
    for i, run in enumerate(run_distances):
        t = np.float32(np.sort(np.asarray(dataset_distances[i], np.float32))[count - 1]) + np.float32(1e-3)
        recalls.append(float(sum(1 for dd in list(run)[:count] if np.float32(dd) <= t)))
    recalls_arr = np.asarray(recalls, np.float32)
    mean = float(recalls_arr.sum() / (len(recalls) * count)) if len(recalls) else 0.0
    std = float(recalls_arr.std() / count) if len(recalls) else 0.0
    return mean, std, recalls


class RunMetrics:
    """The run / query / cluster granularities of the reference's RunMetrics (src/utils/metrics/mod.rs:14-35,116-262) for one
    batch: what the reference writes to sqlite (result_schema.sql: search_metrics, search_metrics_query,
    search_metrics_cluster) as plain rows. The cluster rows come from the device's per-visit log (CLANN_X_VISIT_LOG:
    cluster_n_candidates = heap adds that returned true, cluster_distance_computations incl. the prune-test evaluation,
    cluster time = nanoseconds of the visit on the device); per-query wall time is the batch time divided by the batch
    (queries run concurrently)."""

    def __init__(self, config: "Config", dataset_len: int, total_search_time_s: float, counters: dict,
                 run_distances=None, ground_truth_distances=None, indexing_duration_s: float = 0.0, visit_log=None):
        self.visit_log = None if visit_log is None else np.asarray(visit_log, np.uint32)
        self.config, self.dataset_len = config, int(dataset_len)
        self.total_search_time_s = float(total_search_time_s)
        self.indexing_duration_s = float(indexing_duration_s)
        self.distance_computations = np.asarray(counters["distance_computations"], np.uint64)
        self.candidates = np.asarray(counters["candidates"], np.uint64)
        self.clusters_visited = np.asarray(counters["clusters_visited"], np.uint32)
        nq = len(self.distance_computations)
        self.queries_per_second = (nq / self.total_search_time_s) if self.total_search_time_s > 0 else 0.0  # mod.rs:261
        self.recall_mean = self.recall_std = 0.0
        self.recalls: List[float] = []
        if ground_truth_distances is not None and run_distances is not None:
            self.recall_mean, self.recall_std, self.recalls = get_recall_values(ground_truth_distances, run_distances, int(config.k))

    def run_row(self) -> dict:
        c = self.config
        return dict(num_clusters_factor=float(c.num_clusters_factor), num_tables=int(c.num_tables), k=int(c.k), delta=float(c.delta),
                    dataset=c.dataset_name, dataset_len=self.dataset_len, indexing_duration_s=self.indexing_duration_s,
                    total_search_time_s=self.total_search_time_s, queries_per_second=self.queries_per_second,
                    recall_mean=self.recall_mean, recall_std=self.recall_std)

    def query_rows(self) -> List[dict]:
        nq = len(self.distance_computations)
        per_query_s = self.total_search_time_s / nq if nq else 0.0
        return [dict(query_idx=i, query_time_s=per_query_s, distance_computations=int(self.distance_computations[i]),
                     n_candidates=int(self.candidates[i]), clusters_visited=int(self.clusters_visited[i]),
                     recall=(self.recalls[i] / float(self.config.k)) if self.recalls else None) for i in range(nq)]

    def cluster_rows(self) -> List[dict]:
        """search_metrics_cluster (result_schema.sql:73-90, written by sqlite.rs:248-283): one row per visited cluster of every
        query; cluster_idx is the visit's ordinal as in the reference's enumerate(), `cluster` the cluster's id (extra)."""
        if self.visit_log is None:
            raise ValueError("no per-visit log: run with granularity='cluster'")
        rows = []
        for q, per_query in enumerate(self.visit_log):
            for v, (c1, added, dc, ns) in enumerate(per_query):
                if c1 == 0:
                    break
                rows.append(dict(query_idx=q, cluster_idx=v, cluster=int(c1) - 1, n_candidates=int(added),
                                 cluster_time_s=float(ns) * 1e-9, cluster_distance_computations=int(dc)))
        return rows

    def to_json(self, granularity: str = "query") -> str:
        import json
        out = {"run": self.run_row()}
        if granularity in ("query", "cluster"):
            out["queries"] = self.query_rows()
        if granularity == "cluster":
            out["clusters"] = self.cluster_rows()
        return json.dumps(out)

    def write_csv(self, path: str) -> None:
        rows = self.query_rows()
        with open(path, "w") as f:
            f.write("# " + ",".join(f"{k}={v}" for k, v in self.run_row().items()) + "\n")
            f.write("query_idx,query_time_s,distance_computations,n_candidates,clusters_visited,recall\n")
            for r in rows:
                f.write(",".join("" if r[k] is None else str(r[k]) for k in
                                 ("query_idx", "query_time_s", "distance_computations", "n_candidates", "clusters_visited", "recall")) + "\n")


def run_with_metrics(index: "ClusteredIndex", queries, ground_truth_distances=None, granularity: str = "query",
                     max_visits: int = 64) -> Tuple[Tuple[np.ndarray, np.ndarray, np.ndarray], RunMetrics]:
    """One batch through clann_search with the reference's run metrics filled in (wall clock around the call, as
    benches/distance_benches.rs:57-74 does around its query loop). granularity = "cluster" (MetricsGranularity::Cluster,
    config.rs:9-13) also records the per-visit rows, up to max_visits per query."""
    import time
    if granularity not in ("run", "query", "cluster"):
        raise ValueError("granularity must be 'run', 'query' or 'cluster'")
    index.set_option("visit_log", max_visits if granularity == "cluster" else 0)
    try:
        t0 = time.perf_counter()
        ids, dists, counts = index.search_batch(queries)
        dt = time.perf_counter() - t0
        ctr = index.counters(len(counts))
        log = index.visit_log(len(counts)) if granularity == "cluster" else None
    finally:
        index.set_option("visit_log", 0)
    run = [dists[i, :counts[i]].tolist() for i in range(len(counts))] if ground_truth_distances is not None else None
    return (ids, dists, counts), RunMetrics(index.config, index.data.num_points(), dt, ctr, run, ground_truth_distances, visit_log=log)


def generate_random_unit_vectors(n: int, dimensions: int, seed: Optional[int] = None) -> np.ndarray:
    """src/utils/mod.rs:101-114 — U[0,1) entries, normalised."""
    rng = np.random.default_rng(seed)
    v = rng.random((n, dimensions), dtype=np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def brute_force_search(data: np.ndarray, query: np.ndarray, k: int) -> np.ndarray:
    """src/utils/mod.rs:116-131 — exact neighbours for measuring recall (host-side measurement helper)."""
    data = np.asarray(data, np.float32)
    query = np.asarray(query, np.float32)
    d = 1.0 - (data @ query) / (np.linalg.norm(data, axis=1) * np.linalg.norm(query))
    return np.argsort(d, kind="stable")[:k].astype(np.uint32)
