"""TEST INFRASTRUCTURE — ctypes front-ends for the two CPU checkers.

* ``OracleLib``  -> oracle/liboracle.so        (plain-C restatement, clann_oracle.c)
* ``RefLib``     -> oracle/_ref/libpuffinn_ref.so (the real reference PUFFINN, ref_binder.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
Nothing under clann_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpuffinn_ref.so")

_vp, _u32, _u64, _i32, _f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (always) and _ref (only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OrcFunctions(C.Structure):
    _fields_ = [
        ("d", _u32), ("sl", _u32), ("m", _u32), ("bpf", _u32), ("fph", _u32), ("cut", _u32), ("L", _u32),
        ("rotations", _u32), ("planes", _vp), ("signs", _vp), ("est", _vp), ("eps", _f32),
    ]


class OrcIndex(C.Structure):
    _fields_ = [
        ("fn", OrcFunctions), ("n", _u32), ("q15", _vp), ("sketches", _vp), ("hashes", _vp), ("indices", _vp),
        ("prefix_index", _vp),
    ]


class OrcTrace(C.Structure):
    _fields_ = [
        ("distance_computations", _u32), ("candidates", _u32), ("stop_depth", _u32), ("stop_table", _u32),
        ("n_batches", _u32), ("kth", _f32), ("max_sketch_diff", _u32), ("passing", _vp), ("passing_cap", _u64),
        ("passing_len", _u64), ("batch_sizes", _vp), ("batch_cap", _u64),
    ]


class Functions:
    """A PUFFINN function set held as numpy arrays (planes Q15, FHT signs, collision estimates)."""

    def __init__(self, d, L, planes, signs, est, eps=np.float32(0.005), rotations=3):
        self.d = int(d)
        self.sl = (self.d + 15) // 16 * 16
        m, p = 0, 1
        while p < self.d:
            m, p = m + 1, p * 2
        self.m = m
        self.bpf = m + 1
        self.fph = (24 + self.bpf - 1) // self.bpf
        self.cut = self.bpf * self.fph - 24
        self.L = int(L)
        self.rotations = int(rotations)
        self.planes = np.ascontiguousarray(planes, dtype=np.int16).reshape(2048, self.sl)
        self.signs = np.ascontiguousarray(signs, dtype=np.int8).reshape(self.L * self.fph, self.rotations << m)
        self.est = np.ascontiguousarray(est, dtype=np.float32).reshape(m + 2, 201)
        self.eps = np.float32(eps)

    def c_struct(self) -> OrcFunctions:
        return OrcFunctions(self.d, self.sl, self.m, self.bpf, self.fph, self.cut, self.L, self.rotations,
                            _ptr(self.planes), _ptr(self.signs), _ptr(self.est), float(self.eps))


class OracleIndex:
    def __init__(self, lib: "OracleLib", handle):
        self.lib, self.h = lib, handle
        s = C.cast(handle, C.POINTER(OrcIndex)).contents
        self.n, self.L, self.d, self.sl = s.n, s.fn.L, s.fn.d, s.fn.sl
        self.m, self.bpf, self.fph, self.cut = s.fn.m, s.fn.bpf, s.fn.fph, s.fn.cut
        self._s = s

    def _arr(self, ptr, shape, dtype):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    @property
    def q15(self):
        return self._arr(self._s.q15, (self.n, self.sl), np.int16)

    @property
    def sketches(self):
        return self._arr(self._s.sketches, (self.n, 32), np.uint64)

    @property
    def hashes(self):
        return self._arr(self._s.hashes, (self.L, self.n + 24), np.uint32)

    @property
    def indices(self):
        return self._arr(self._s.indices, (self.L, self.n + 24), np.uint32)

    def functions(self) -> Functions:
        f = self._s.fn
        planes = self._arr(f.planes, (2048, f.sl), np.int16).copy()
        signs = self._arr(f.signs, (f.L * f.fph, f.rotations << f.m), np.int8).copy()
        est = self._arr(f.est, (f.m + 2, 201), np.float32).copy()
        return Functions(f.d, f.L, planes, signs, est, f.eps, f.rotations)

    def codes(self, q15row):
        out = np.zeros(self.L, np.uint32)
        q = np.ascontiguousarray(q15row, np.int16)
        self.lib.lib.orc_codes(C.byref(self._s.fn), _ptr(q), _ptr(out))
        return out

    def sketch(self, q15row):
        out = np.zeros(32, np.uint64)
        q = np.ascontiguousarray(q15row, np.int16)
        self.lib.lib.orc_sketch(C.byref(self._s.fn), _ptr(q), _ptr(out))
        return out

    def query_ranges(self, codes):
        codes = np.ascontiguousarray(codes, np.uint32)
        anchors = np.zeros(self.L, np.uint32)
        ranges = np.zeros((24, self.L, 2), np.uint32)
        self.lib.lib.orc_query_ranges(self.h, _ptr(codes), _ptr(anchors), _ptr(ranges))
        return anchors, ranges

    def search(self, q, k, recall, max_sim=float("-inf"), trace=False, filter_type=0):
        """filter_type: 0 = FilterType::Default, 1 = None, 2 = Simple (collection.hpp:22-34)."""
        q = np.ascontiguousarray(q, np.float32)
        out = np.zeros(max(k, 1), np.uint32)
        tr = OrcTrace()
        passing = batches = None
        if trace:
            passing = np.zeros(1 << 20, np.uint32)
            batches = np.zeros(1 << 14, np.uint32)
            tr.passing, tr.passing_cap = _ptr(passing), passing.size
            tr.batch_sizes, tr.batch_cap = _ptr(batches), batches.size
        if filter_type:
            cnt = self.lib.lib.orc_index_search_filter(self.h, _ptr(q), k, float(recall), float(max_sim), int(filter_type), _ptr(out),
                                                       C.byref(tr))
        else:
            cnt = self.lib.lib.orc_index_search(self.h, _ptr(q), k, float(recall), float(max_sim), _ptr(out), C.byref(tr))
        info = dict(distance_computations=tr.distance_computations, candidates=tr.candidates, stop_depth=tr.stop_depth,
                    stop_table=tr.stop_table, n_batches=tr.n_batches, kth=tr.kth, max_sketch_diff=tr.max_sketch_diff)
        if trace:
            info["passing"] = passing[: tr.passing_len].copy()
            info["batch_sizes"] = batches[: min(tr.n_batches, batches.size)].copy()
        return out[:cnt].copy(), info

    def failure_probability(self, depth, tables, max_tables, kth):
        return self.lib.lib.orc_failure_probability(C.byref(self._s.fn), depth, tables, max_tables, float(kth))

    def free(self):
        if self.h:
            self.lib.lib.orc_index_free(self.h)
            self.h = None


class OracleLib:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.orc_to_q15.restype, L.orc_to_q15.argtypes = C.c_int16, [_f32]
        L.orc_from_q15.restype, L.orc_from_q15.argtypes = _f32, [C.c_int16]
        L.orc_storage_len.restype, L.orc_storage_len.argtypes = _u32, [_u32]
        L.orc_ceil_log.restype, L.orc_ceil_log.argtypes = _u32, [_u32]
        L.orc_store_q15.argtypes = [_vp, _u32, _u32, _vp]
        L.orc_dot_i16.restype, L.orc_dot_i16.argtypes = C.c_int16, [_vp, _vp, _u32]
        L.orc_similarity.restype, L.orc_similarity.argtypes = _f32, [_vp, _vp, _u32]
        L.orc_fht.argtypes = [_vp, _u32]
        L.orc_fht_hash.restype, L.orc_fht_hash.argtypes = _u32, [_vp, _u32, _vp]
        L.orc_codes.argtypes = [_vp, _vp, _vp]
        L.orc_sketch.argtypes = [_vp, _vp, _vp]
        L.orc_sort_pairs_24.argtypes = [_vp, _vp, _u32, _vp, _vp]
        L.orc_failure_probability.restype = _f32
        L.orc_failure_probability.argtypes = [_vp, _u32, _u64, _u64, _f32]
        L.orc_max_sketch_diff.restype, L.orc_max_sketch_diff.argtypes = _u32, [_f32]
        L.orc_maxbuffer_run.restype = _i32
        L.orc_maxbuffer_run.argtypes = [_u32, _vp, _vp, _i32, _vp, _vp, _vp]
        L.orc_index_build.restype, L.orc_index_build.argtypes = _vp, [_vp, _vp, _u32]
        L.orc_index_import.restype, L.orc_index_import.argtypes = _vp, [_vp, _u64]
        L.orc_index_free.argtypes = [_vp]
        L.orc_anchor.restype, L.orc_anchor.argtypes = _u32, [_vp, _u32, _u32]
        L.orc_query_ranges.argtypes = [_vp, _vp, _vp, _vp]
        L.orc_index_search.restype = _i32
        L.orc_index_search.argtypes = [_vp, _vp, _u32, _f32, _f32, _vp, _vp]
        L.orc_index_search_filter.restype = _i32
        L.orc_index_search_filter.argtypes = [_vp, _vp, _u32, _f32, _f32, _i32, _vp, _vp]
        L.orc_num_clusters.restype, L.orc_num_clusters.argtypes = _u64, [_f32, _u64]
        L.orc_ndarray_dot.restype, L.orc_ndarray_dot.argtypes = _f32, [_vp, _vp, C.c_size_t]
        L.orc_distance_point.restype, L.orc_distance_point.argtypes = _f32, [_vp, _f32, _vp, _u32]
        L.orc_gmm.restype, L.orc_gmm.argtypes = _u64, [_vp, _u64, _u32, _u64, _vp, _vp, _vp]
        L.orc_topk_run.restype = _i32
        L.orc_topk_run.argtypes = [_u32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]
        L.orc_sort_clusters.argtypes = [_vp, _u32, _vp, _u64, _vp, _vp]
        L.orc_clann_create.restype = _vp
        L.orc_clann_create.argtypes = [_vp, _u64, _u32, _u32, _f32, _u64, _vp, _vp, _vp]
        L.orc_clann_set_cluster_stream.restype = _i32
        L.orc_clann_set_cluster_stream.argtypes = [_vp, _u64, _vp, _u64]
        L.orc_clann_build_cluster.restype, L.orc_clann_build_cluster.argtypes = _i32, [_vp, _u64, _vp]
        L.orc_clann_search.restype, L.orc_clann_search.argtypes = _i32, [_vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_clann_search_visits.restype, L.orc_clann_search_visits.argtypes = _i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64]
        L.orc_clann_free.argtypes = [_vp]

    # L0 helpers -----------------------------------------------------------------------------------
    def store_q15(self, v, sl=None):
        v = np.ascontiguousarray(v, np.float32)
        d = v.shape[-1]
        sl = sl or (d + 15) // 16 * 16
        flat = v.reshape(-1, d)
        out = np.zeros((flat.shape[0], sl), np.int16)
        for i in range(flat.shape[0]):
            self.lib.orc_store_q15(flat[i].ctypes.data, d, sl, out[i].ctypes.data)
        return out.reshape(v.shape[:-1] + (sl,))

    def dot_i16(self, a, b):
        a, b = np.ascontiguousarray(a, np.int16), np.ascontiguousarray(b, np.int16)
        return int(self.lib.orc_dot_i16(_ptr(a), _ptr(b), a.size))

    def similarity(self, a, b):
        a, b = np.ascontiguousarray(a, np.int16), np.ascontiguousarray(b, np.int16)
        return float(self.lib.orc_similarity(_ptr(a), _ptr(b), a.size))

    def fht(self, buf, m):
        buf = np.ascontiguousarray(buf, np.float32).copy()
        self.lib.orc_fht(_ptr(buf), m)
        return buf

    def sort_pairs(self, hashes, idx):
        hashes, idx = np.ascontiguousarray(hashes, np.uint32), np.ascontiguousarray(idx, np.uint32)
        ho, io = np.zeros_like(hashes), np.zeros_like(idx)
        self.lib.orc_sort_pairs_24(_ptr(hashes), _ptr(idx), hashes.size, _ptr(ho), _ptr(io))
        return ho, io

    def maxbuffer_run(self, k, ids, vals):
        ids, vals = np.ascontiguousarray(ids, np.uint32), np.ascontiguousarray(vals, np.float32)
        oi, ov = np.zeros(max(k, 1), np.uint32), np.zeros(max(k, 1), np.float32)
        mv = _f32(0)
        c = self.lib.orc_maxbuffer_run(k, _ptr(ids), _ptr(vals), ids.size, _ptr(oi), _ptr(ov), C.byref(mv))
        return oi[:c], ov[:c], mv.value

    # L1 ------------------------------------------------------------------------------------------
    def index_build(self, fn: Functions, data) -> OracleIndex:
        data = np.ascontiguousarray(data, np.float32)
        cs = fn.c_struct()
        return OracleIndex(self, self.lib.orc_index_build(C.byref(cs), _ptr(data), data.shape[0]))

    def index_import(self, stream: bytes) -> OracleIndex:
        buf = np.frombuffer(stream, np.uint8)
        h = self.lib.orc_index_import(_ptr(buf), buf.size)
        if not h:
            raise ValueError("oracle could not parse the Index::serialize stream")
        return OracleIndex(self, h)

    # L3 ------------------------------------------------------------------------------------------
    def num_clusters(self, factor, n):
        return int(self.lib.orc_num_clusters(float(np.float32(factor)), n))

    def gmm(self, data, K):
        data = np.ascontiguousarray(data, np.float32)
        n, d = data.shape
        kk = min(n, K) if n <= K else K
        centers, assign, radii = np.zeros(max(K, n), np.uint64), np.zeros(n, np.uint64), np.zeros(max(K, n), np.float32)
        kk = self.lib.orc_gmm(_ptr(data), n, d, K, _ptr(centers), _ptr(assign), _ptr(radii))
        return centers[:kk].copy(), assign, radii[:kk].copy()

    def distance_point(self, row, q):
        row, q = np.ascontiguousarray(row, np.float32), np.ascontiguousarray(q, np.float32)
        norm = np.float32(np.sqrt(np.float32(self.lib.orc_ndarray_dot(_ptr(row), _ptr(row), row.size))))
        return float(self.lib.orc_distance_point(_ptr(row), float(norm), _ptr(q), row.size))

    def topk_run(self, k, dists, ids):
        """TopKClosestHeap (heap.rs): returns (to_list as [(dist, id)], added flags, get_top as (id, dist) or None)."""
        dists, ids = np.ascontiguousarray(dists, np.float32), np.ascontiguousarray(ids, np.uint64)
        od, oi = np.zeros(max(k, 1), np.float32), np.zeros(max(k, 1), np.uint64)
        added = np.zeros(max(len(ids), 1), np.uint8)
        tid, td = _u64(0), _f32(0)
        c = self.lib.orc_topk_run(k, _ptr(dists), _ptr(ids), len(ids), _ptr(od), _ptr(oi), _ptr(added), C.byref(tid), C.byref(td))
        top = (int(tid.value), float(td.value)) if c > 0 else None
        return [(float(od[i]), int(oi[i])) for i in range(c)], added[: len(ids)].astype(bool), top

    def sort_clusters(self, data, centers, q):
        data, q = np.ascontiguousarray(data, np.float32), np.ascontiguousarray(q, np.float32)
        centers = np.ascontiguousarray(centers, np.uint64)
        order = np.zeros(centers.size, np.uint64)
        self.lib.orc_sort_clusters(_ptr(data), data.shape[1], _ptr(centers), centers.size, _ptr(q), _ptr(order))
        return order

    def clann(self, data, k, delta, centers, assignment, radii) -> "OracleClann":
        return OracleClann(self, data, k, delta, centers, assignment, radii)


class OracleClann:
    def __init__(self, lib: OracleLib, data, k, delta, centers, assignment, radii):
        self.lib = lib
        self.data = np.ascontiguousarray(data, np.float32)
        self.k = int(k)
        self.centers = np.ascontiguousarray(centers, np.uint64)
        self.assignment = np.ascontiguousarray(assignment, np.uint64)
        self.radii = np.ascontiguousarray(radii, np.float32)
        self.K = self.centers.size
        n, d = self.data.shape
        self.h = lib.lib.orc_clann_create(_ptr(self.data), n, d, self.k, float(np.float32(delta)), self.K,
                                          _ptr(self.centers), _ptr(self.assignment), _ptr(self.radii))

    def set_cluster_stream(self, ci, stream: bytes):
        buf = np.frombuffer(stream, np.uint8)
        rc = self.lib.lib.orc_clann_set_cluster_stream(self.h, ci, _ptr(buf), buf.size)
        if rc != 0:
            raise ValueError(f"cluster {ci}: stream rejected ({rc})")

    def build_cluster(self, ci, fn: Functions):
        cs = fn.c_struct()
        return self.lib.lib.orc_clann_build_cluster(self.h, ci, C.byref(cs))

    def search(self, q):
        q = np.ascontiguousarray(q, np.float32)
        ids, dists = np.zeros(max(self.k, 1), np.uint64), np.zeros(max(self.k, 1), np.float32)
        order, counters = np.zeros(self.K, np.uint64), np.zeros(3, np.uint64)
        c = self.lib.lib.orc_clann_search(self.h, _ptr(q), _ptr(ids), _ptr(dists), _ptr(order), _ptr(counters))
        if c < 0:
            raise RuntimeError(f"oracle clann search failed ({c})")
        return ids[:c].copy(), dists[:c].copy(), order, dict(visited=int(counters[0]), distance_computations=int(counters[1]),
                                                            candidates=int(counters[2]))

    def search_visits(self, q, cap=64):
        """search() plus the per-visit rows of the reference's cluster-granularity metrics: uint64 [visits, 3] =
        (cluster, points_added, cluster_distance_computations)."""
        q = np.ascontiguousarray(q, np.float32)
        ids, dists = np.zeros(max(self.k, 1), np.uint64), np.zeros(max(self.k, 1), np.float32)
        order, counters = np.zeros(self.K, np.uint64), np.zeros(3, np.uint64)
        log = np.zeros((cap, 3), np.uint64)
        c = self.lib.lib.orc_clann_search_visits(self.h, _ptr(q), _ptr(ids), _ptr(dists), _ptr(order), _ptr(counters), _ptr(log), cap)
        if c < 0:
            raise RuntimeError(f"oracle clann search failed ({c})")
        return ids[:c].copy(), dists[:c].copy(), log[: min(int(counters[0]), cap)].copy()

    def free(self):
        if self.h:
            self.lib.lib.orc_clann_free(self.h)
            self.h = None


class RefIndex:
    def __init__(self, lib: "RefLib", h, d):
        self.lib, self.h, self.d = lib, h, d

    @property
    def n(self):
        return self.lib.lib.ref_index_size(self.h)

    @property
    def sl(self):
        return self.lib.lib.ref_index_storage_len(self.h)

    def insert(self, rows):
        rows = np.ascontiguousarray(rows, np.float32).reshape(-1, self.d)
        for r in rows:
            if self.lib.lib.ref_index_insert(self.h, r.ctypes.data, self.d) != 0:
                raise RuntimeError("insert failed")

    def rebuild(self, L):
        return self.lib.lib.ref_index_rebuild(self.h, L)

    def search(self, q, k, recall, max_sim=float("-inf"), filter_type=0):
        """filter_type: the reference's FilterType (collection.hpp:22-34): 0 = Default, 1 = None, 2 = Simple."""
        q = np.ascontiguousarray(q, np.float32)
        out, met = np.zeros(max(k, 1), np.uint32), np.zeros(4, np.uint32)
        if filter_type:
            c = self.lib.lib.ref_index_search_filter(self.h, _ptr(q), k, float(recall), float(max_sim), int(filter_type), _ptr(out),
                                                     out.size, _ptr(met))
        else:
            c = self.lib.lib.ref_index_search(self.h, _ptr(q), k, float(recall), float(max_sim), _ptr(out), out.size, _ptr(met))
        if c < 0:
            raise RuntimeError("reference search threw")
        return out[:c].copy(), dict(distance_computations=int(met[0]), candidates=int(met[1]), hash_length=int(met[2]),
                                    considered_maps=int(met[3]))

    def serialize(self) -> bytes:
        size = self.lib.lib.ref_index_serialize(self.h, None, 0)
        buf = np.zeros(size, np.uint8)
        self.lib.lib.ref_index_serialize(self.h, _ptr(buf), size)
        return buf.tobytes()

    def point(self, i):
        out = np.zeros(self.sl, np.int16)
        self.lib.lib.ref_dump_point(self.h, i, _ptr(out))
        return out

    def store_q15(self, v):
        v = np.ascontiguousarray(v, np.float32)
        out = np.zeros(self.sl, np.int16)
        self.lib.lib.ref_store_q15(self.h, _ptr(v), _ptr(out))
        return out

    def query_codes(self, q, L):
        q = np.ascontiguousarray(q, np.float32)
        out = np.zeros(L, np.uint64)
        c = self.lib.lib.ref_query_codes(self.h, _ptr(q), _ptr(out))
        assert c == L
        return out

    def query_sketches(self, q):
        q = np.ascontiguousarray(q, np.float32)
        out = np.zeros(32, np.uint64)
        self.lib.lib.ref_query_sketches(self.h, _ptr(q), _ptr(out))
        return out

    def point_sketches(self, i):
        out = np.zeros(32, np.uint64)
        self.lib.lib.ref_point_sketches(self.h, i, _ptr(out))
        return out

    def table(self, t):
        ln = self.lib.lib.ref_table(self.h, t, None, None)
        h, i = np.zeros(ln, np.uint32), np.zeros(ln, np.uint32)
        self.lib.lib.ref_table(self.h, t, _ptr(h), _ptr(i))
        return h, i

    def query_ranges(self, q, L):
        q = np.ascontiguousarray(q, np.float32)
        anchors, ranges = np.zeros(L, np.uint32), np.zeros((24, L, 2), np.uint32)
        c = self.lib.lib.ref_query_ranges(self.h, _ptr(q), _ptr(anchors), _ptr(ranges))
        assert c == L
        return anchors, ranges

    def similarity(self, q, i):
        q = np.ascontiguousarray(q, np.float32)
        return float(self.lib.lib.ref_similarity(self.h, _ptr(q), i))

    def failure_probability(self, depth, tables, max_tables, sim):
        return float(self.lib.lib.ref_failure_probability(self.h, depth, tables, max_tables, float(sim)))

    def max_sketch_diff(self, sim):
        return int(self.lib.lib.ref_max_sketch_diff(self.h, float(sim)))

    def free(self):
        if self.h:
            self.lib.lib.ref_index_free(self.h)
            self.h = None


class RefLib:
    """The real reference, when oracle/_ref/libpuffinn_ref.so exists (built here, shipped to the GPU box)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self, path: str = REF_SO):
        L = self.lib = C.CDLL(path)
        L.ref_seed.argtypes = [_u64]
        L.ref_index_create.restype, L.ref_index_create.argtypes = _vp, [_i32]
        L.ref_index_free.argtypes = [_vp]
        L.ref_index_insert.restype, L.ref_index_insert.argtypes = _i32, [_vp, _vp, _i32]
        L.ref_index_rebuild.restype, L.ref_index_rebuild.argtypes = _u64, [_vp, C.c_uint]
        L.ref_index_search.restype = _i32
        L.ref_index_search.argtypes = [_vp, _vp, C.c_uint, _f32, _f32, _vp, _i32, _vp]
        L.ref_index_search_filter.restype = _i32
        L.ref_index_search_filter.argtypes = [_vp, _vp, C.c_uint, _f32, _f32, _i32, _vp, _i32, _vp]
        L.ref_index_serialize.restype, L.ref_index_serialize.argtypes = _u64, [_vp, _vp, _u64]
        L.ref_index_deserialize.restype, L.ref_index_deserialize.argtypes = _vp, [_vp, _u64]
        L.ref_index_size.restype, L.ref_index_size.argtypes = _u32, [_vp]
        L.ref_index_storage_len.restype, L.ref_index_storage_len.argtypes = _u32, [_vp]
        L.ref_dump_point.argtypes = [_vp, _u32, _vp]
        L.ref_store_q15.argtypes = [_vp, _vp, _vp]
        L.ref_query_codes.restype, L.ref_query_codes.argtypes = _i32, [_vp, _vp, _vp]
        L.ref_query_sketches.argtypes = [_vp, _vp, _vp]
        L.ref_point_sketches.argtypes = [_vp, _u32, _vp]
        L.ref_table.restype, L.ref_table.argtypes = _u64, [_vp, _u32, _vp, _vp]
        L.ref_query_ranges.restype, L.ref_query_ranges.argtypes = _i32, [_vp, _vp, _vp, _vp]
        L.ref_similarity.restype, L.ref_similarity.argtypes = _f32, [_vp, _vp, _u32]
        L.ref_failure_probability.restype = _f32
        L.ref_failure_probability.argtypes = [_vp, C.c_uint, C.c_uint, C.c_uint, _f32]
        L.ref_max_sketch_diff.restype, L.ref_max_sketch_diff.argtypes = C.c_uint, [_vp, _f32]
        L.ref_to_q15.restype, L.ref_to_q15.argtypes = C.c_int16, [_f32]
        L.ref_from_q15.restype, L.ref_from_q15.argtypes = _f32, [C.c_int16]
        L.ref_dot_i16.restype, L.ref_dot_i16.argtypes = C.c_int16, [_vp, _vp, C.c_uint]
        L.ref_dot_i16_simple.restype, L.ref_dot_i16_simple.argtypes = C.c_int16, [_vp, _vp, C.c_uint]
        L.ref_fht.argtypes = [_vp, _i32]
        L.ref_maxbuffer.restype = _i32
        L.ref_maxbuffer.argtypes = [C.c_uint, _vp, _vp, _i32, _vp, _vp, _vp]
        L.ref_clann_num_clusters.restype, L.ref_clann_num_clusters.argtypes = _u64, [_f32, _u64]
        L.ref_clann_gmm.restype, L.ref_clann_gmm.argtypes = _u64, [_vp, _u64, _u32, _u64, _vp, _vp, _vp]
        L.ref_clann_create.restype = _vp
        L.ref_clann_create.argtypes = [_vp, _u64, _u32, _u32, _u32, _f32, _u64, _vp, _vp, _vp, _u64]
        L.ref_clann_build_all.argtypes = [_vp]
        L.ref_clann_build_cluster.argtypes = [_vp, _u64]
        L.ref_clann_build_seconds.restype, L.ref_clann_build_seconds.argtypes = C.c_double, [_vp]
        L.ref_clann_cluster_serialize.restype = _u64
        L.ref_clann_cluster_serialize.argtypes = [_vp, _u64, _vp, _u64]
        L.ref_clann_search.restype = _i32
        L.ref_clann_search.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
        L.ref_clann_last_counters.argtypes = [_vp, _vp, _vp, _vp]
        L.ref_clann_free.argtypes = [_vp]
        L.ref_omp_threads.restype = _i32

    def seed(self, s):
        self.lib.ref_seed(int(s))

    def index(self, d, rows=None, L=None, seed=None) -> RefIndex:
        if seed is not None:
            self.seed(seed)
        h = self.lib.ref_index_create(d)
        if not h:
            raise RuntimeError("reference index_create failed")
        ix = RefIndex(self, h, d)
        if rows is not None:
            ix.insert(rows)
            if L is not None and ix.rebuild(L) == 0:
                raise RuntimeError("reference rebuild failed")
        return ix

    def index_from_stream(self, stream: bytes) -> RefIndex:
        """puffinn::Index(std::istream&) (collection.hpp:147-170) over a serialized stream."""
        buf = np.frombuffer(stream, np.uint8)
        self.lib.ref_index_deserialize.restype = _vp
        self.lib.ref_index_deserialize.argtypes = [_vp, _u64]
        h = self.lib.ref_index_deserialize(_ptr(buf), buf.size)
        if not h:
            raise RuntimeError("the reference could not deserialize the stream")
        d = int(np.frombuffer(stream[:4], np.uint32)[0])
        return RefIndex(self, h, d)

    def fht(self, buf, m):
        buf = np.ascontiguousarray(buf, np.float32).copy()
        self.lib.ref_fht(_ptr(buf), m)
        return buf

    def dot_i16(self, a, b, simple=False):
        a, b = np.ascontiguousarray(a, np.int16), np.ascontiguousarray(b, np.int16)
        f = self.lib.ref_dot_i16_simple if simple else self.lib.ref_dot_i16
        return int(f(_ptr(a), _ptr(b), a.size))

    def maxbuffer(self, k, ids, vals):
        ids, vals = np.ascontiguousarray(ids, np.uint32), np.ascontiguousarray(vals, np.float32)
        oi, ov = np.zeros(max(k, 1), np.uint32), np.zeros(max(k, 1), np.float32)
        mv = _f32(0)
        c = self.lib.ref_maxbuffer(k, _ptr(ids), _ptr(vals), ids.size, _ptr(oi), _ptr(ov), C.byref(mv))
        return oi[:c], ov[:c], mv.value

    def gmm(self, data, K):
        data = np.ascontiguousarray(data, np.float32)
        n, d = data.shape
        centers, assign, radii = np.zeros(max(K, n), np.uint64), np.zeros(n, np.uint64), np.zeros(max(K, n), np.float32)
        kk = self.lib.ref_clann_gmm(_ptr(data), n, d, K, _ptr(centers), _ptr(assign), _ptr(radii))
        return centers[:kk].copy(), assign, radii[:kk].copy()

    def clann(self, data, L, k, delta, centers, assignment, radii, seed_base=1234) -> "RefClann":
        return RefClann(self, data, L, k, delta, centers, assignment, radii, seed_base)


class RefClann:
    """CLANN search loop (restated) over *reference* PUFFINN indices; indices are built lazily per visited cluster."""

    def __init__(self, lib: RefLib, data, L, k, delta, centers, assignment, radii, seed_base):
        self.lib = lib
        self.data = np.ascontiguousarray(data, np.float32)
        self.k = int(k)
        self.centers = np.ascontiguousarray(centers, np.uint64)
        self.assignment = np.ascontiguousarray(assignment, np.uint64)
        self.radii = np.ascontiguousarray(radii, np.float32)
        self.K = self.centers.size
        n, d = self.data.shape
        self.h = lib.lib.ref_clann_create(_ptr(self.data), n, d, L, self.k, float(np.float32(delta)), self.K,
                                          _ptr(self.centers), _ptr(self.assignment), _ptr(self.radii), seed_base)
        self.search_seconds = C.c_double(0)

    def build_all(self):
        self.lib.lib.ref_clann_build_all(self.h)

    def build_cluster(self, ci):
        self.lib.lib.ref_clann_build_cluster(self.h, ci)

    @property
    def build_seconds(self):
        return self.lib.lib.ref_clann_build_seconds(self.h)

    def cluster_stream(self, ci) -> bytes:
        size = self.lib.lib.ref_clann_cluster_serialize(self.h, ci, None, 0)
        if size == 0:
            return b""
        buf = np.zeros(size, np.uint8)
        self.lib.lib.ref_clann_cluster_serialize(self.h, ci, _ptr(buf), size)
        return buf.tobytes()

    def search(self, q):
        q = np.ascontiguousarray(q, np.float32)
        ids, dists = np.zeros(max(self.k, 1), np.uint64), np.zeros(max(self.k, 1), np.float32)
        order = np.zeros(self.K, np.uint64)
        c = self.lib.lib.ref_clann_search(self.h, _ptr(q), _ptr(ids), _ptr(dists), _ptr(order), C.byref(self.search_seconds))
        v, dc, cd = _u64(0), _u64(0), _u64(0)
        self.lib.lib.ref_clann_last_counters(self.h, C.byref(v), C.byref(dc), C.byref(cd))
        return ids[:c].copy(), dists[:c].copy(), order, dict(visited=v.value, distance_computations=dc.value, candidates=cd.value)

    def free(self):
        if self.h:
            self.lib.lib.ref_clann_free(self.h)
            self.h = None
