"""More GPU tests through the C ABI: golden fixtures, the legacy CPUFFINN_* symbols, edge cases the reference's own
tests and code paths cover (tiny / ragged inputs, brute-force clusters, n <= K, zero vectors, k larger than a cluster),
size-independent properties at a larger size, and the cluster-sharded search emulated with two shards on one GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _single_cluster_index(cb, data, L, k, delta, stream):
    ix = cb.init_with_config(data, cb.Config(L, 1.0, k, delta, "golden"))
    n = len(data)
    ix.set_clustering([0], np.zeros(n, np.uint64), [2.0])
    ix.import_reference(0, stream)
    ix.build()
    return ix


@pytest.mark.parametrize("name", ["puffinn_d25", "puffinn_d100"])
def test_gpu_matches_golden_fixture(name):
    """One PUFFINN index = a CLANN index with one imposed cluster: build bytes, query codes / sketches, result sets and
    counters must equal what the real reference produced (tests/golden/make_golden.py)."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    L, data, queries = int(g["L"]), g["data"], g["queries"]
    n = len(data)
    for si, (k, rec, ms) in enumerate(g["searches"]):
        if not np.isinf(ms):
            continue  # an explicit max_sim only exists in the legacy call: test_gpu_matches_golden_fixture_with_max_sim
        ix = _single_cluster_index(cb, data, L, int(k), float(rec), g["stream"].tobytes())
        ids, dists, counts = ix.search_batch(queries)
        ctr = ix.counters(len(queries))
        codes = ix.export(cl.X_QUERY_CODES, 0, np.uint32).reshape(len(queries), L)
        sks = ix.export(cl.X_QUERY_SKETCHES, 0, np.uint64).reshape(len(queries), 32)
        assert np.array_equal(codes, g["query_codes"]) and np.array_equal(sks, g["query_sketches"])
        for qi in range(len(queries)):
            cnt = int(g["res_cnt"][si, qi])
            assert counts[qi] == cnt
            assert sorted(ids[qi, :cnt].tolist()) == sorted(g["res_ids"][si, qi, :cnt].tolist())
            assert int(ctr["distance_computations"][qi]) == int(g["res_met"][si, qi, 0])
            assert int(ctr["candidates"][qi]) == int(g["res_met"][si, qi, 1])
        # the stored tables themselves
        th = ix.export(cl.X_TABLE_HASHES, 0, np.uint32).reshape(L, n)
        key = th.astype(np.int64)
        assert np.all(np.diff(key, axis=1) >= 0)
        ix.close()


def _write_record_file(path, name, payload):
    """One CLB2REC record (the flat container CPUFFINN_save_index writes; layout in index.cu)."""
    with open(path, "wb") as f:
        f.write(b"CLB2REC\0")
        f.write(np.array([len(name), 0], np.uint32).tobytes())
        f.write(np.array([len(payload)], np.uint64).tobytes())
        f.write(name.encode())
        f.write(payload)


@pytest.mark.parametrize("name", ["puffinn_d25", "puffinn_d100"])
def test_gpu_matches_golden_fixture_with_max_sim(name, tmp_path):
    """Every golden search, including those with an explicit max_sim (CLANN's addition to the stop rule, collection.hpp:324-330,
    927-943), through the legacy symbol that takes it: CPUFFINN_load_from_file on the reference's own stream, then
    CPUFFINN_search_cosine(query, k, recall, max_sim). Ids (as sets; the golden order is the reference's tie order) and
    the distance-computation counter must equal the real reference's."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    path = str(tmp_path / "golden.clb2")
    _write_record_file(path, "index_0", g["stream"].tobytes())
    L = cl.load()
    handle = L.CPUFFINN_load_from_file(path.encode(), b"index_0")
    assert handle, cl.last_error()
    d = g["queries"].shape[1]
    checked_ms = 0
    for si, (k, rec, ms) in enumerate(g["searches"]):
        k = int(k)
        for qi, q in enumerate(g["queries"]):
            q = np.ascontiguousarray(q, np.float32)
            ptr = L.CPUFFINN_search_cosine(handle, q.ctypes.data, k, float(rec), float(ms), d)
            assert ptr
            got = [ptr[i] for i in range(max(k, 1))]
            C.CDLL(None).free(ptr)
            cnt = int(g["res_cnt"][si, qi])
            assert sorted(got[:cnt]) == sorted(g["res_ids"][si, qi, :cnt].tolist()), (si, qi)
            assert all(v == 0xFFFFFFFF for v in got[cnt:])
            assert L.CPUFFINN_get_distance_computations() == int(g["res_met"][si, qi, 0]), (si, qi)
        checked_ms += int(not np.isinf(ms))
    assert checked_ms > 0  # the fixture does hold searches with a finite max_sim


def test_legacy_save_refuses_foreign_container(tmp_path):
    """ClusteredIndex::serialize in the unmodified crate creates an HDF5 file and hands the same path to CPUFFINN_save_index
    (index.rs:526-552). Without HDF5 here the call must not append records to such a file: it refuses and leaves the file
    untouched; a fresh or CLB2REC file is appended to as before."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    data = util.planted(400, 16, 55, n_centers=2)
    index, _ = cb.PuffinnIndex.new(cb.AngularData(data), 6)
    foreign = tmp_path / "index_x.h5"
    foreign.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    before = foreign.read_bytes()
    cl.load().CPUFFINN_save_index(index.raw, str(foreign).encode(), 3)
    assert foreign.read_bytes() == before
    assert "not a libclann_b200 record file" in cl.last_error()
    assert not cl.load().CPUFFINN_load_from_file(str(foreign).encode(), b"index_3")
    ok = tmp_path / "fresh.clb2"
    index.save_to_file(str(ok), 3)
    assert ok.read_bytes()[:8] == b"CLB2REC\0"


def test_legacy_abi_recall():
    """src/puffinn_binds/puffinn.rs:179-226 — n=1000, d=25, L=40, k in {1,10}, recall in {0.2,0.5,0.95}, 100 queries:
    hits >= 0.8 * recall * k * 100."""
    import clann_b200 as cb
    rng = np.random.default_rng(42)
    data = util.uniform_sphere(1000, 25, 1)
    index, mem = cb.PuffinnIndex.new(cb.AngularData(data), 40)
    assert mem > 0
    queries = util.uniform_sphere(100, 25, 2)
    ex = util.exact_distances(data, queries)
    for k in (1, 10):
        truth = np.argsort(ex, axis=1)[:, :k]
        for recall in (0.2, 0.5, 0.95):
            hits = 0
            for qi, q in enumerate(queries):
                res = index.search(q, k, float("inf"), recall)
                assert len(res) == k
                hits += len(set(res) & set(truth[qi].tolist()))
                assert cb.api.get_distance_computations() > 0
            assert hits >= 0.8 * recall * k * 100, (k, recall, hits)


def test_legacy_abi_small_index_is_exact_and_padded():
    """Fewer than 100 points: Q15 brute force (collection.hpp:550-555); fewer results than k are 0xFFFFFFFF-padded
    (hardening deviation documented in include/clann_b200.h)."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    data = util.uniform_sphere(40, 16, 3)
    index, _ = cb.PuffinnIndex.new(cb.AngularData(data), 4)
    q = data[7] + 0.01
    res = index.search(q, 5, float("inf"), 0.9)
    truth = np.argsort(util.exact_distances(data, q[None])[0])[:5]
    assert res[0] == 7 and set(res) == set(truth.tolist())
    L = cl.load()
    ptr = L.CPUFFINN_search_cosine(index.raw, np.ascontiguousarray(q, np.float32).ctypes.data, 64, 0.9, float("-inf"), 16)
    got = [ptr[i] for i in range(64)]
    C.CDLL(None).free(ptr)
    assert sorted(got[:40]) == list(range(40)) and all(v == 0xFFFFFFFF for v in got[40:])
    # wrong dimension -> NULL, never a crash
    assert not L.CPUFFINN_search_cosine(index.raw, np.zeros(8, np.float32).ctypes.data, 3, 0.9, 0.0, 8)


def test_edge_cases_against_oracle(oracle):
    """n <= K (every point its own centre, gmm.rs:26-31), all-brute-force clusters, zero vector, duplicate points,
    k larger than every cluster."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    rng = np.random.default_rng(9)
    # (a) n <= K
    data = util.uniform_sphere(5, 8, 1)
    ix = cb.init_with_config(data, cb.Config(4, 10.0, 3, 0.9))
    ix.build()
    assert ix.num_clusters == 5
    assert np.array_equal(ix.export(cl.X_CENTERS, 0, np.uint64), np.arange(5, dtype=np.uint64))
    cen, asg, rad = (ix.export(cl.X_CENTERS, 0, np.uint64), ix.export(cl.X_ASSIGNMENT, 0, np.uint64), ix.export(cl.X_RADII, 0, np.float32))
    orc = oracle.clann(data, 3, 0.9, cen, asg, rad)
    for q in [data[2], util.uniform_sphere(1, 8, 2)[0]]:
        res = ix.search(q)
        o_ids, o_d, _, _ = orc.search(q)
        # the prune test uses the heap top even while the heap holds fewer than k entries (index.rs:342-361, heap.rs:38-40):
        # an exact hit (distance 0, radius 0) ends the search with a single result
        assert [i for _, i in res] == [int(x) for x in o_ids] and np.array_equal(np.float32([dd for dd, _ in res]), o_d)
    assert ix.search(data[2]) == [(0.0, 2)]
    orc.free()
    ix.close()
    # (b) small clusters only (all brute force) + zero vector + duplicates; compare with the oracle's CLANN loop
    data = util.planted(300, 12, 5, n_centers=6)
    data[10] = 0.0
    data[20] = data[21]
    ix = cb.init_with_config(data, cb.Config(6, 0.4, 4, 0.9))
    ix.build()
    K = ix.num_clusters
    cen, asg, rad = (ix.export(cl.X_CENTERS, 0, np.uint64), ix.export(cl.X_ASSIGNMENT, 0, np.uint64), ix.export(cl.X_RADII, 0, np.float32))
    oc, oa, orad = oracle.gmm(data, K)
    assert np.array_equal(cen, oc) and np.array_equal(asg, oa)
    brute = ix.export(cl.X_BRUTE, 0, np.uint8)
    orc = oracle.clann(data, 4, 0.9, cen, asg, rad)
    assert brute.any()                     # the shape does produce brute-force clusters (index.rs:204-205)
    for ci in range(K):                    # ... and whatever is not brute force is searched by the oracle over the same tables
        if not brute[ci]:
            orc.set_cluster_stream(ci, ix.export(cl.X_REFERENCE_STREAM, ci, np.uint8).tobytes())
    qs = np.concatenate([util.planted_queries(data, 20, 6), data[20:22], data[10:11] + 1e-3])
    ix.set_option("visit_log", 8)
    ids, dists, counts = ix.search_batch(qs)
    ctr = ix.counters(len(qs))
    vlog = ix.visit_log(len(qs))
    for i, q in enumerate(qs):
        o_ids, o_d, _, o_ctr = orc.search(q)
        assert util.same_ids_up_to_ties(ids[i, : counts[i]], dists[i, : counts[i]], o_ids.astype(np.uint32), o_d), i
        assert int(ctr["clusters_visited"][i]) == o_ctr["visited"] and int(ctr["distance_computations"][i]) == o_ctr["distance_computations"]
        # per-visit metric rows incl. brute-force visits (index.rs:364-378: list length as distance computations)
        _, _, o_log = orc.search_visits(q, 8)
        assert np.array_equal(vlog[i][: len(o_log), 0].astype(np.int64) - 1, o_log[:, 0].astype(np.int64)), i
        assert np.array_equal(vlog[i][: len(o_log), 1:3].astype(np.uint64), o_log[:, 1:3]), (i, vlog[i], o_log)
    ix.close(); orc.free()
    # (c) k larger than every cluster -> every cluster is brute force (index.rs:204-205), results exact
    data = util.planted(1200, 10, 7)
    ix = cb.init_with_config(data, cb.Config(4, 0.4, 500, 0.9))
    ix.build()
    assert ix.export(cl.X_BRUTE, 0, np.uint8).all()
    q = util.planted_queries(data, 4, 8)
    ids, dists, counts = ix.search_batch(q)
    assert (counts == 500).all() and np.all(np.diff(dists, axis=1) >= 0)
    ix.close()


def test_not_built_and_dimension_errors():
    import clann_b200 as cb
    data = util.uniform_sphere(200, 8, 1)
    ix = cb.init_with_config(data, cb.Config(4, 0.4, 3, 0.9))
    with pytest.raises(cb.IndexNotFound):
        ix.search(data[0])
    ix.build()
    with pytest.raises(cb.PuffinnSearchError):
        ix.search(np.zeros(9, np.float32))
    ids, dists, counts = ix.search_batch(np.zeros((0, 8), np.float32))
    assert ids.shape == (0, 3)
    ix.close()


@pytest.fixture(scope="module")
def big():
    import clann_b200 as cb
    data = util.planted(200_000, 96, 21)
    ix = cb.init_with_config(data, cb.Config(84, 0.4, 10, 0.9, "big"))
    ix.set_option("seed", 5)
    ix.build()
    return data, ix


def test_properties_at_scale(big):
    """Size-independent properties on 200k x 96: tables sorted by (hash, id), every local id present exactly once per
    table, perm is a permutation, recall@10 >= 0.9 (utils/mod.rs:59-95), determinism."""
    from clann_b200 import _lib as cl
    data, ix = big
    n = len(data)
    perm = ix.export(cl.X_PERM, 0, np.uint32)
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    off = ix.export(cl.X_OFFSETS, 0, np.uint64)
    asg = ix.export(cl.X_ASSIGNMENT, 0, np.uint64)
    brute = ix.export(cl.X_BRUTE, 0, np.uint8)
    for ci in np.flatnonzero(brute == 0)[:5]:
        nc = int(off[ci + 1] - off[ci])
        members = perm[int(off[ci]): int(off[ci + 1])]
        assert np.all(np.diff(members.astype(np.int64)) > 0) and np.all(asg[members] == ci)  # index.rs:188-192
        th = ix.export(cl.X_TABLE_HASHES, int(ci), np.uint32).reshape(84, nc)
        ti = ix.export(cl.X_TABLE_INDICES, int(ci), np.uint32).reshape(84, nc)
        key = (th.astype(np.uint64) << np.uint64(32)) | ti
        assert np.all(np.diff(key.astype(np.int64), axis=1) > 0) and th.max() < (1 << 24)
        assert np.array_equal(np.sort(ti, axis=1), np.broadcast_to(np.arange(nc, dtype=np.uint32), (84, nc)))
    q = util.planted_queries(data, 1000, 22)
    ids, dists, counts = ix.search_batch(q)
    assert util.recall_at_k(data, q, dists, counts, 10) >= 0.9
    ids2, dists2, counts2 = ix.search_batch(q)
    assert np.array_equal(ids, ids2) and np.array_equal(dists.view(np.uint32), dists2.view(np.uint32))
    # returned distances are the fp32 cosine distances of the returned ids, ascending, no duplicates
    ex = util.exact_distances(data, q[:50])
    for i in range(50):
        c = counts[i]
        assert np.allclose(dists[i, :c], ex[i, ids[i, :c]], atol=2e-6) and np.all(np.diff(dists[i, :c]) >= 0)
        assert len(set(ids[i, :c].tolist())) == c


def test_sharded_search_equals_single_gpu(big):
    """Two shards emulated on one GPU (kernels never wait on each other, so sequential launches are safe): stepping with
    a state exchange between steps reproduces the unsharded search exactly, and each shard only probes its own clusters."""
    import torch

    import clann_b200 as cb
    from clann_b200 import _lib as cl
    from clann_b200.distributed import _DeviceBytes
    data, full = big
    L = cl.load()
    q = util.planted_queries(data, 600, 23)
    q = np.concatenate([q, util.uniform_sphere(8, 96, 24)])  # a few queries that wander across many clusters
    ids0, d0, c0 = full.search_batch(q)
    shards = []
    for r in range(2):
        ix = cb.init_with_config(data, cb.Config(84, 0.4, 10, 0.9, "big"))
        ix.set_option("seed", 5)
        ix.set_option("shard_count", 2)
        ix.set_option("shard_rank", r)
        ix.build()
        shards.append(ix)
    dev = torch.device("cuda", 0)
    dq = torch.from_numpy(q).to(dev)
    nq = len(q)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sb = int(L.clann_state_bytes(shards[0].handle))
    views = []
    for ix in shards:
        assert L.clann_search_begin(ix.handle, dq.data_ptr(), nq, stream) == 0
        views.append(torch.as_tensor(_DeviceBytes(L.clann_state_ptr(ix.handle), nq * sb), device=dev))
    steps = 0
    while True:
        for ix in shards:
            assert L.clann_search_step(ix.handle, stream) == 0
        torch.cuda.synchronize()
        allst = torch.cat(views).contiguous()
        active = [C.c_uint64(0), C.c_uint64(0)]
        for r, ix in enumerate(shards):
            assert L.clann_search_merge(ix.handle, allst.data_ptr(), 2, C.byref(active[r]), stream) == 0
        steps += 1
        assert active[0].value == active[1].value
        if active[0].value == 0:
            break
        assert steps < 1000
    for ix in shards:
        ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
        dd = torch.empty((nq, 10), dtype=torch.float32, device=dev)
        cc = torch.empty(nq, dtype=torch.int32, device=dev)
        assert L.clann_search_end(ix.handle, ids.data_ptr(), dd.data_ptr(), cc.data_ptr(), stream) == 0
        torch.cuda.synchronize()
        assert np.array_equal(ids.cpu().numpy().view(np.uint32), ids0)
        assert np.array_equal(dd.cpu().numpy().view(np.uint32), d0.view(np.uint32))
        assert np.array_equal(cc.cpu().numpy().view(np.uint32), c0)
    assert steps >= 2  # at least one hand-over happened
    for ix in shards:
        ix.close()


def test_bucket_directory_matches_tables(big):
    """The bucket directory (role of PrefixMap::prefix_index, prefixmap.hpp:86,231-240, at 12 bits): entry b of table t is the
    lower bound of (b << 12) in that table's sorted codes; entry 4096 is the cluster size."""
    from clann_b200 import _lib as cl
    _, ix = big
    off = ix.export(cl.X_OFFSETS, 0, np.uint64)
    brute = ix.export(cl.X_BRUTE, 0, np.uint8)
    keys = (np.arange(4097, dtype=np.uint32) << np.uint32(12))
    for ci in np.flatnonzero(brute == 0)[:4]:
        nc = int(off[ci + 1] - off[ci])
        th = ix.export(cl.X_TABLE_HASHES, int(ci), np.uint32).reshape(84, nc)
        dr = ix.export(cl.X_TABLE_DIR, int(ci), np.uint32).reshape(84, 4097)
        for t in range(84):
            assert np.array_equal(dr[t], np.searchsorted(th[t], keys, side="left").astype(np.uint32))


_VARIANT_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from tests import util
import clann_b200 as cb
data = util.planted(60_000, 100, 31)
q = np.concatenate([util.planted_queries(data, 400, 32), util.uniform_sphere(40, 100, 33)])
ix = cb.init_with_config(data, cb.Config(84, 0.4, 10, 0.9, "variant"))
ix.set_option("seed", 77)
ix.build()
ids, dists, counts = ix.search_batch(q)
ctr = ix.counters(len(q))
np.savez(sys.argv[2], ids=ids, dists=dists, counts=counts, cand=ctr["candidates"], dc=ctr["distance_computations"],
         vis=ctr["clusters_visited"])
"""


def test_probe_kernel_variants_agree(tmp_path):
    """Probe schedule variants (clann_tune knobs: dense first-visit similarities on / off, first-visit anchors / ranges /
    stream on / off, memo in shared or global memory or absent, launch shapes): every variant must return identical ids, distance bits and reference counters (candidates, distance_computations, clusters visited) for the same
    index and queries — planted queries plus uniform ones that walk many clusters."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    variants = (("warp", {}), ("warp_no_dense_sims", {"CLANN_TUNE_DENSE_SIMS": "0"}),
                ("no_tensor_centre_scoring", {"CLANN_TUNE_TC_CENTER": "0"}), ("no_tensor_sketches", {"CLANN_TUNE_TC_SKETCH": "0"}),
                ("no_tensor_at_all", {"CLANN_TUNE_TC_CENTER": "0", "CLANN_TUNE_TC_SKETCH": "0"}),
                ("first_stream", {"CLANN_TUNE_FIRST_STREAM": "1"}),
                ("first_stream_cta", {"CLANN_TUNE_FIRST_STREAM": "2"}),
                ("first_stream_cta_cut_short", {"CLANN_TUNE_FIRST_STREAM": "2", "CLANN_TUNE_FIRST_STREAM_CAP": "320"}),
                ("first_stream_cut_short", {"CLANN_TUNE_FIRST_STREAM": "1", "CLANN_TUNE_FIRST_STREAM_CAP": "256"}),   # most visits outlive it
                ("first_stream_tiny", {"CLANN_TUNE_FIRST_STREAM": "1", "CLANN_TUNE_FIRST_STREAM_CAP": "64"}),
                ("no_first_ranges_overlap", {"CLANN_TUNE_FIRST_RANGES_OVERLAP": "0"}),
                ("warp_no_first_ranges", {"CLANN_TUNE_FIRST_RANGES": "0"}),
                ("warp_all_first_ranges", {"CLANN_TUNE_FIRST_RANGES": "1"}), ("warp_no_smem_memo", {"CLANN_TUNE_PROBE_SMEM_MEMO": "0"}), ("warp_nomemo", {"CLANN_TUNE_PROBE_NOMEMO": "1"}),
                ("warp_small_grid", {"CLANN_TUNE_PROBE_WARPS": "4", "CLANN_TUNE_PROBE_CTAS": "1"}),
                ("warp_longest_first_prefetch", {"CLANN_TUNE_ORDER_LONGEST_FIRST": "1", "CLANN_TUNE_PROBE_PREFETCH_ROWS": "1"}))
    for name, env in variants:
        out = str(tmp_path / (name + ".npz"))
        e = {k: v for k, v in os.environ.items() if not k.startswith("CLANN_TUNE_")}
        e.update(env)
        subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT, root, out], check=True, env=e, timeout=600)
        outs[name] = np.load(out)
    ref = outs["warp"]
    assert ref["vis"].max() > 3  # the uniform queries do walk several clusters
    for name, _ in variants[1:]:
        o = outs[name]
        for key in ("ids", "counts", "cand", "dc", "vis"):
            assert np.array_equal(ref[key], o[key]), (name, key)
        assert np.array_equal(ref["dists"].view(np.uint32), o["dists"].view(np.uint32)), name


def test_legacy_save_and_load_round_trip(tmp_path, reflib):
    """CPUFFINN_save_index / CPUFFINN_load_from_file (c_binder.cpp:4-36,106-146, puffinn.rs:61-75,121-141): the saved
    record is the reference's own serialization (the real reference loads it), and an index loaded from it answers every
    query exactly like the index that was saved — without a rebuild."""
    import clann_b200 as cb
    data = util.planted(1500, 25, 51, n_centers=3)
    index, _ = cb.PuffinnIndex.new(cb.AngularData(data), 30)
    path = str(tmp_path / "saved.clb2")
    index.save_to_file(path, 7)
    index.save_to_file(path, 8)  # records are appended (the reference appends datasets to one file)
    loaded = cb.PuffinnIndex.new_from_file(path, 8)
    with pytest.raises(cb.api.SerializeError):
        cb.PuffinnIndex.new_from_file(path, 9)
    queries = util.planted_queries(data, 40, 52)
    raw = open(path, "rb").read()
    name_len = int(np.frombuffer(raw[8:12], np.uint32)[0])
    payload_len = int(np.frombuffer(raw[16:24], np.uint64)[0])
    assert raw[:8] == b"CLB2REC\0" and raw[24:24 + name_len] == b"index_7"
    ref = reflib.index_from_stream(raw[24 + name_len: 24 + name_len + payload_len])
    for q in queries:
        a = index.search(q, 10, float("inf"), 0.9)
        dc = cb.api.get_distance_computations()
        b = loaded.search(q, 10, float("inf"), 0.9)
        assert a == b and cb.api.get_distance_computations() == dc
        rids, met = ref.search(q, 10, 0.9)
        assert sorted(rids.tolist()) == sorted(a) and met["distance_computations"] == dc


def test_async_batches_equal_stream_ordered_calls():
    """clann_search_device_async / clann_search_flush (batch pipelining on two internal streams with their own workspaces):
    six batches in flight back to back — different queries, one of a different size — return exactly what the
    stream-ordered host call returns for each of them."""
    import torch
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    data = util.planted(40_000, 64, 61)
    ix = cb.init_with_config(data, cb.Config(40, 0.4, 10, 0.9, "async"))
    ix.set_option("seed", 9)
    ix.build()
    batches = [util.planted_queries(data, 700 if i != 3 else 333, 70 + i) for i in range(6)]
    batches[4] = util.uniform_sphere(700, 64, 99)   # walks many clusters
    expected = [ix.search_batch(q) for q in batches]
    dev = torch.device("cuda", 0)
    d_q = [torch.from_numpy(q).to(dev) for q in batches]
    outs = [(torch.empty((len(q), 10), dtype=torch.int32, device=dev), torch.empty((len(q), 10), dtype=torch.float32, device=dev),
             torch.empty(len(q), dtype=torch.int32, device=dev)) for q in batches]
    torch.cuda.synchronize()
    lib = cl.load()
    for q, o in zip(d_q, outs):
        assert lib.clann_search_device_async(ix.handle, q.data_ptr(), q.shape[0], o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) == 0, cl.last_error()
    assert lib.clann_search_flush(ix.handle, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    for (ids, dists, counts), o in zip(expected, outs):
        assert np.array_equal(ids.view(np.uint32), o[0].cpu().numpy().view(np.uint32))
        assert np.array_equal(dists.view(np.uint32), o[1].cpu().numpy().view(np.uint32))
        assert np.array_equal(counts.view(np.uint32), o[2].cpu().numpy().view(np.uint32))


def test_async_host_batches_equal_synchronous_call():
    """clann_search_async / clann_search_wait: host buffers in and out (pinned), copies and search of every batch on its
    internal stream; five batches back to back return what clann_search returns."""
    import torch
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    data = util.planted(30_000, 100, 63)
    ix = cb.init_with_config(data, cb.Config(30, 0.4, 10, 0.9, "async-host"))
    ix.set_option("seed", 10)
    ix.build()
    batches = [util.planted_queries(data, 500, 80 + i) for i in range(5)]
    expected = [ix.search_batch(q) for q in batches]
    h_q = [torch.from_numpy(q).pin_memory() for q in batches]
    outs = [(torch.empty((500, 10), dtype=torch.int32).pin_memory(), torch.empty((500, 10), dtype=torch.float32).pin_memory(),
             torch.empty(500, dtype=torch.int32).pin_memory()) for _ in batches]
    lib = cl.load()
    for q, o in zip(h_q, outs):
        assert lib.clann_search_async(ix.handle, q.data_ptr(), 500, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) == 0, cl.last_error()
    assert lib.clann_search_wait(ix.handle) == 0
    for (ids, dists, counts), o in zip(expected, outs):
        assert np.array_equal(ids.view(np.uint32), o[0].numpy().view(np.uint32))
        assert np.array_equal(dists.view(np.uint32), o[1].numpy().view(np.uint32))
        assert np.array_equal(counts.view(np.uint32), o[2].numpy().view(np.uint32))


def test_properties_at_full_glove100_shape():
    """BASELINE.json configs[2] at full size (1 183 514 x 100, L = 84, K = 435): size-independent properties — perm is a
    permutation, sampled tables are sorted by (hash, id) and hold every local id once, recall@10 >= 0.9 against exact fp32
    neighbours (utils/mod.rs:59-95), the dense first-visit path and the gather path agree, two batches in flight agree with
    the blocking call, determinism."""
    import torch
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    n, d = 1_183_514, 100
    data = util.planted(n, d, 42)
    ix = cb.init_with_config(data, cb.Config(84, 0.4, 10, 0.9, "glove-100-shape"))
    ix.set_option("seed", 1234)
    ix.build()
    assert ix.num_clusters == 435
    perm = ix.export(cl.X_PERM, 0, np.uint32)
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    off = ix.export(cl.X_OFFSETS, 0, np.uint64)
    brute = ix.export(cl.X_BRUTE, 0, np.uint8)
    for ci in np.flatnonzero(brute == 0)[[0, 200, -1]]:
        nc = int(off[ci + 1] - off[ci])
        th = ix.export(cl.X_TABLE_HASHES, int(ci), np.uint32).reshape(84, nc)
        ti = ix.export(cl.X_TABLE_INDICES, int(ci), np.uint32).reshape(84, nc)
        key = (th.astype(np.uint64) << np.uint64(32)) | ti
        assert np.all(np.diff(key.astype(np.int64), axis=1) > 0) and th.max() < (1 << 24)
        assert np.array_equal(np.sort(ti, axis=1), np.broadcast_to(np.arange(nc, dtype=np.uint32), (84, nc)))
    q = util.planted_queries(data, 2000, 43)
    ids, dists, counts = ix.search_batch(q)
    ctr = ix.counters(len(q))
    # recall against exact neighbours computed on the device (torch is test plumbing here)
    dd = torch.from_numpy(data).cuda()
    qq = torch.from_numpy(q).cuda()
    kth = torch.cat([torch.topk(qq[s:s + 200] @ dd.T, 10, dim=1).values[:, 9] for s in range(0, len(q), 200)])
    kth = (1.0 - kth).cpu().numpy()
    hit = sum(int(np.sum(dists[i, :counts[i]] <= kth[i] + 1e-3)) for i in range(len(q)))
    assert hit / (len(q) * 10) >= 0.9
    del dd, qq
    # the gather path (no dense similarities) returns the same ids, distances and counters
    cl.tune("dense_sims", 0)
    try:
        ids2, dists2, counts2 = ix.search_batch(q)
        ctr2 = ix.counters(len(q))
    finally:
        cl.tune("dense_sims", 1)
    assert np.array_equal(ids, ids2) and np.array_equal(dists.view(np.uint32), dists2.view(np.uint32)) and np.array_equal(counts, counts2)
    for key in ("candidates", "distance_computations", "clusters_visited"):
        assert np.array_equal(ctr[key], ctr2[key]), key
    ids3, dists3, _ = ix.search_batch(q)
    assert np.array_equal(ids, ids3) and np.array_equal(dists.view(np.uint32), dists3.view(np.uint32))


def test_serialize_and_init_from_file_round_trip(tmp_path):
    """lib.rs:41-47,255-264 (index.rs:107-162,511-557): serialize writes `config`, `clusters` and one `index_{i}` record per
    PUFFINN cluster; init_from_file(data, path) answers every query exactly like the index that was saved — ids, distance
    bits, counters — and the per-cluster streams are the reference's bytes (the real reference loads one of them)."""
    import json
    import clann_b200 as cb
    from clann_b200 import api
    data = util.planted(20_000, 48, 71)
    ix = cb.init_with_config(data, cb.Config(24, 0.4, 10, 0.9, "roundtrip"))
    ix.set_option("seed", 5)
    ix.build()
    q = np.concatenate([util.planted_queries(data, 300, 72), util.uniform_sphere(20, 48, 73)])
    ids, dists, counts = ix.search_batch(q)
    ctr = ix.counters(len(q))
    with pytest.raises(cb.api.SerializeError):
        cb.serialize(ix, str(tmp_path / "no_such_dir"))
    path = cb.serialize(ix, str(tmp_path))
    assert os.path.basename(path) == "index_roundtrip_k0.40_L24.clb2"
    rec = api._read_records(path)
    cfg = json.loads(rec["config"])
    clusters = json.loads(rec["clusters"])
    assert cfg["num_tables"] == 24 and cfg["k"] == 10 and len(clusters) == ix.num_clusters
    assert sorted(sum((c["assignment"] for c in clusters), [])) == list(range(len(data)))
    assert all((f"index_{c['idx']}" in rec) == (not c["brute_force"]) for c in clusters)
    loaded = cb.init_from_file(data, path)
    ids2, dists2, counts2 = loaded.search_batch(q)
    ctr2 = loaded.counters(len(q))
    assert np.array_equal(ids, ids2) and np.array_equal(dists.view(np.uint32), dists2.view(np.uint32)) and np.array_equal(counts, counts2)
    for key in ("candidates", "distance_computations", "clusters_visited"):
        assert np.array_equal(ctr[key], ctr2[key]), key
    # a second save of the loaded index writes the same records
    os.makedirs(tmp_path / "again")
    rec2 = api._read_records(cb.serialize(loaded, str(tmp_path / "again")))
    assert rec2.keys() == rec.keys() and all(rec2[k] == rec[k] for k in rec if k != "config") and json.loads(rec2["config"]) == cfg


def _lib_x():
    from clann_b200 import _lib
    return _lib


def test_run_with_metrics_on_device():
    """run_with_metrics (the run / query rows of the reference's RunMetrics): counters per query come from the device, recall
    from get_recall_values against exact distances, queries/s from the wall clock around the batched call."""
    import clann_b200 as cb
    data = util.planted(15_000, 32, 81)
    ix = cb.init_with_config(data, cb.Config(30, 0.4, 10, 0.9, "metrics"))
    ix.build()
    q = util.planted_queries(data, 200, 82)
    gt = np.sort(util.exact_distances(data, q), axis=1)[:, :10]
    (ids, dists, counts), m = cb.run_with_metrics(ix, q, gt)
    assert m.queries_per_second > 0 and m.recall_mean >= 0.9 and len(m.query_rows()) == 200
    ctr = ix.counters(200)
    rows = m.query_rows()
    assert [r["distance_computations"] for r in rows] == ctr["distance_computations"].tolist()
    assert [r["n_candidates"] for r in rows] == ctr["candidates"].tolist()
    assert m.run_row()["dataset_len"] == 15_000 and m.run_row()["dataset"] == "metrics"
    # MetricsGranularity::Cluster: one row per visited cluster; results and counters are those of the plain call
    (ids2, dists2, counts2), mc = cb.run_with_metrics(ix, q, gt, granularity="cluster")
    assert np.array_equal(ids, ids2) and np.array_equal(dists.view(np.uint32), dists2.view(np.uint32))
    crow = mc.cluster_rows()
    vis = ix.counters(200)["clusters_visited"]
    assert len(crow) == int(vis.sum())
    per_q = np.zeros(200, np.int64)
    dc_q = np.zeros(200, np.int64)
    for r in crow:
        assert r["cluster_idx"] == per_q[r["query_idx"]] and 0 <= r["cluster"] < ix.num_clusters and r["cluster_time_s"] > 0
        per_q[r["query_idx"]] += 1
        dc_q[r["query_idx"]] += r["cluster_distance_computations"]
    assert np.array_equal(per_q, vis)
    brute = ix.export(_lib_x().X_BRUTE, 0, np.uint8)
    if not brute.any():  # PUFFINN's counter + one prune-test evaluation for every visit after the first (index.rs:348,421)
        assert np.array_equal(dc_q, ctr["distance_computations"].astype(np.int64) + vis - 1)
    import json
    assert len(json.loads(mc.to_json("cluster"))["clusters"]) == len(crow)


@pytest.mark.parametrize("shape", [(30_000, 64, "planted"), (12_000, 100, "uniform"), (9_000, 36, "planted")])
def test_gmm_kernel_variants_match_oracle(oracle, shape):
    """greedy_minimum_maximum (gmm.rs:21-62) on the device in its three forms — 8 lanes per row, vectorised two lanes per row,
    vectorised with the triangle-inequality filter that skips rows — against the oracle's restatement: centres, assignment and
    radii bit for bit (the filter must never change a decision, on clustered and on unclustered data)."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    n, d, kind = shape
    data = util.planted(n, d, 91) if kind == "planted" else util.uniform_sphere(n, d, 92)
    data[17] = 0.0                      # a zero vector: NaN distances must stay on the evaluated path
    data[40] = data[41]                 # duplicates: zero residual distance
    K = oracle.num_clusters(0.4, n)
    oc, oa, orad = oracle.gmm(data, K)
    for knobs in ({"gmm_vec": 0}, {"gmm_prune": 0}, {"assign_partition": 0}, {}):
        for k_, v in knobs.items():
            cl.tune(k_, v)
        try:
            ix = cb.init_with_config(data, cb.Config(4, 0.4, 5, 0.9, "gmm"))
            ix.build()
            cen = ix.export(cl.X_CENTERS, 0, np.uint64)
            asg = ix.export(cl.X_ASSIGNMENT, 0, np.uint64)
            rad = ix.export(cl.X_RADII, 0, np.float32)
            # member lists (index.rs:188-192): ascending point order inside every cluster, clusters back to back — by the
            # chunked counting sort (default) and by the one-segment radix sort (assign_partition = 0)
            perm = ix.export(cl.X_PERM, 0, np.uint32)
            offs = ix.export(cl.X_OFFSETS, 0, np.uint64)
            assert np.array_equal(perm, np.argsort(asg, kind="stable").astype(np.uint32)), knobs
            assert np.array_equal(offs, np.concatenate([[0], np.cumsum(np.bincount(asg.astype(np.int64), minlength=K))]).astype(np.uint64))
            ix.close()
        finally:
            for k_ in knobs:
                cl.tune(k_, 1)
        assert np.array_equal(cen, oc), knobs
        assert np.array_equal(asg, oa), knobs
        assert np.array_equal(rad.view(np.uint32), orad.view(np.uint32)), knobs


def test_cluster_sharded_search_two_ranks_in_process(big):
    """clann_search_sharded (the multi-GPU mode of SURVEY.md 8e: route by nearest cluster, reference loop on the owner, one
    all-reduce(min) of bounds, pruned second round on every rank, one all-gather + k-way merge) with two ranks as two threads
    of this process on one GPU, talking through clann_set_collectives. Against the single-GPU search of the same index:
    identical results whenever the walk of a query stayed inside one cluster, every rank gets the same answer, every visit of
    the single-GPU search is also made here (visits are a superset), recall is at least the single-GPU recall."""
    import threading
    import torch
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    from clann_b200.distributed import InProcessTransport
    data, full = big
    L = cl.load()
    q = np.concatenate([util.planted_queries(data, 1200, 123), util.uniform_sphere(12, 96, 124)])
    nq, k = len(q), 10
    ids0, d0, c0 = full.search_batch(q)
    ctr0 = full.counters(nq)
    shards = []
    for r in range(2):
        ix = cb.init_with_config(data, cb.Config(84, 0.4, 10, 0.9, "big"))
        ix.set_option("seed", 5)
        ix.set_option("shard_count", 2)
        ix.set_option("shard_rank", r)
        ix.build()
        shards.append(ix)
    # the single-pass entry points refuse a sharded index (it holds tables for its own clusters only)
    with pytest.raises(cb.ConfigError):
        shards[0].search_batch(q[:4])
    dev = torch.device("cuda", 0)
    dq = torch.from_numpy(q).to(dev)
    transport = InProcessTransport(2, dev)
    outs, errors, per_rank = [None, None], [], [None, None]

    def run(rank):
        try:
            torch.cuda.set_device(0)
            ix = shards[rank]
            ag, ar = transport.callbacks(rank)
            assert L.clann_set_collectives(ix.handle, ag, ar, None) == 0
            ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
            dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
            cc = torch.empty(nq, dtype=torch.int32, device=dev)
            for _ in range(2):  # twice: buffers are reused, results must not change
                rc = L.clann_search_sharded(ix.handle, dq.data_ptr(), nq, ids.data_ptr(), dd.data_ptr(), cc.data_ptr(), None)
                assert rc == 0, cl.last_error()
                torch.cuda.synchronize()
            outs[rank] = (ids.cpu().numpy().view(np.uint32), dd.cpu().numpy(), cc.cpu().numpy().view(np.uint32))
            per_rank[rank] = ix.counters(nq)
            # two whole batches in flight (clann_search_sharded_multi): per batch the results of clann_search_sharded
            dq2 = torch.flip(dq, dims=[0]).contiguous()
            o2 = [(torch.empty_like(ids), torch.empty_like(dd), torch.empty_like(cc)) for _ in range(2)]
            arr = lambda ps: (C.c_void_p * 2)(*ps)  # noqa: E731
            rc = L.clann_search_sharded_multi(ix.handle, 2, arr([dq.data_ptr(), dq2.data_ptr()]), nq, arr([o[0].data_ptr() for o in o2]),
                                              arr([o[1].data_ptr() for o in o2]), arr([o[2].data_ptr() for o in o2]), None)
            assert rc == 0, cl.last_error()
            torch.cuda.synchronize()
            assert torch.equal(o2[0][0], ids) and torch.equal(o2[0][1], dd) and torch.equal(o2[0][2], cc)
            assert torch.equal(o2[1][0], torch.flip(ids, dims=[0])) and torch.equal(o2[1][1], torch.flip(dd, dims=[0]))
            assert torch.equal(o2[1][2], torch.flip(cc, dims=[0]))
            # streaming form (clann_search_sharded_submit / _flush): six batches through the four-deep software pipeline, a batch of
            # another size in between; every batch gets the results of the blocking call
            dq3 = dq[: nq // 2].contiguous()
            seq = [dq, dq2, dq3, dq, dq2, dq]
            o3 = [(torch.empty((len(x), k), dtype=torch.int32, device=dev), torch.empty((len(x), k), dtype=torch.float32, device=dev),
                   torch.empty(len(x), dtype=torch.int32, device=dev)) for x in seq]
            for x, o in zip(seq, o3):
                rc = L.clann_search_sharded_submit(ix.handle, x.data_ptr(), len(x), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), None)
                assert rc == 0, cl.last_error()
            # a blocking call while batches are in flight is refused
            assert L.clann_search_sharded(ix.handle, dq.data_ptr(), nq, ids.data_ptr(), dd.data_ptr(), cc.data_ptr(), None) != 0
            assert L.clann_search_sharded_flush(ix.handle, None) == 0, cl.last_error()
            torch.cuda.synchronize()
            want = {id(dq): (ids, dd, cc), id(dq2): tuple(torch.flip(t, dims=[0]) for t in (ids, dd, cc)),
                    id(dq3): (ids[: nq // 2], dd[: nq // 2], cc[: nq // 2])}
            for x, o in zip(seq, o3):
                for got, exp in zip(o, want[id(x)]):
                    assert torch.equal(got, exp)
        except Exception as e:  # noqa: BLE001
            errors.append((rank, repr(e)))
            transport.barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert all(o is not None for o in outs)
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))          # every rank holds the same merged result
    ids, dd, cc = outs[0]
    vis0 = ctr0["clusters_visited"]
    one = vis0 == 1
    assert one.sum() > 800
    assert np.array_equal(ids[one], ids0[one]) and np.array_equal(dd[one].view(np.uint32), d0[one].view(np.uint32)) and np.array_equal(cc[one], c0[one])
    vis = per_rank[0]["clusters_visited"].astype(np.int64) + per_rank[1]["clusters_visited"].astype(np.int64)
    assert np.all(vis >= vis0)                                               # superset of the reference's visits
    assert np.all(np.diff(dd, axis=1)[np.isfinite(dd[:, 1:])] >= 0)          # ascending
    sample = np.arange(0, nq, 4)
    rec_sharded = util.recall_at_k(data, q[sample], dd[sample], cc[sample], k)
    rec_single = util.recall_at_k(data, q[sample], d0[sample], c0[sample], k)
    assert rec_sharded >= rec_single >= 0.9, (rec_sharded, rec_single)
    # the k-th distance never gets worse for the planted queries (same functions, same or more visits)
    worse = np.sum(dd[:1200, k - 1] > d0[:1200, k - 1] + 1e-6)
    assert worse <= 12, worse
    for ix in shards:
        ix.close()


def test_fp16_and_device_resident_ingest():
    """clann_init_with_config_ex (BASELINE.json's fp16 configuration): rows handed over as IEEE half — from host memory or
    already on the device — build exactly the index that the same rows widened to f32 build: identical clustering, tables,
    results, counters."""
    import torch
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    data16 = util.planted(30_000, 96, 131).astype(np.float16)
    wide = data16.astype(np.float32)
    conf = cb.Config(24, 0.4, 20, 0.9, "fp16")
    q = util.planted_queries(wide, 300, 132)
    ref = cb.init_with_config(wide, conf)
    ref.set_option("seed", 3)
    ref.build()
    want = ref.search_batch(q)
    want_ctr = ref.counters(len(q))
    want_tables = ref.export(cl.X_TABLE_HASHES, 1, np.uint32).copy()
    d_rows = torch.from_numpy(data16).cuda()
    for rows, on_device in ((data16, False), (d_rows.data_ptr(), True)):
        ix = cb.ClusteredIndex.from_rows(conf, rows, 30_000, 96, "f16", on_device=on_device)
        ix.set_option("seed", 3)
        ix.build()
        assert np.array_equal(ix.export(cl.X_ASSIGNMENT, 0, np.uint64), ref.export(cl.X_ASSIGNMENT, 0, np.uint64))
        assert np.array_equal(ix.export(cl.X_TABLE_HASHES, 1, np.uint32), want_tables)
        got = ix.search_batch(q)
        ctr = ix.counters(len(q))
        assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(got, want))
        assert all(np.array_equal(ctr[key], want_ctr[key]) for key in ctr)
        ix.close()
    ref.close()


@pytest.mark.parametrize("name", ["puffinn_d25", "puffinn_d100"])
def test_device_collision_estimates_agree_with_reference_table(name):
    """k_cp_trials (the device Monte-Carlo of CrossPolytopeCollisionEstimates, crosspolytope.hpp:16-88: 201 cosine bins x 1000
    repetitions) against the table the real reference drew for the same dimension (golden fixture). Both are 1000-repetition
    estimates, so each entry carries ~1.6 % of sampling noise: like the reference's own statistical test (hash_test.hpp:101-124,
    +-2 % on measured collision rates) the comparison is statistical — mean absolute difference below 2 %, no entry off by more
    than 8 %, every row non-decreasing in the cosine up to noise (4.5 sigma of a difference), P[.][200] ~ 1 and P[0][.] = 1."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    from oracle.pyoracle import OracleLib
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    oi = OracleLib().index_import(g["stream"].tobytes())
    ref_est = oi.functions().est.copy()           # (m + 2) x 201, drawn by the reference
    oi.free()
    data = g["data"]
    ix = cb.init_with_config(data, cb.Config(int(g["L"]), 1.0, 10, 0.9, "est"))
    ix.set_clustering([0], np.zeros(len(data), np.uint64), [2.0])
    ix.set_option("seed", 11)
    ix.build()
    est = ix.export(cl.X_EST, 0, np.float32).reshape(ref_est.shape)
    ix.close()
    assert np.all(est[:, 200] >= 0.99) and np.all(est[0] == 1.0)   # cosine 1 (the reference's own draw reads 0.999 there); zero bits
    diff = np.abs(est - ref_est)
    assert diff.mean() < 0.02, diff.mean()
    assert diff.max() < 0.08, diff.max()
    # adjacent entries are independent 1000-repetition estimates: their difference has sd <= 0.0224, so -0.10 is 4.5 sigma
    assert np.diff(est, axis=1).min() > -0.10, np.diff(est, axis=1).min()
    assert np.all((est >= 0) & (est <= 1))


def test_device_drawn_functions_use_every_bit():
    """The reference's own checks of a hash source and of the filter (hash_source_test.hpp:13-45,76-91, filterer_test.hpp:44-70,
    hash_test.hpp:40-60), on a stand-alone device build (functions drawn by this library, not imported): every table code is below
    2^24, every one of the 24 code bits and of the 32 x 64 sketch bits takes both values over random rows, and the 256 outcomes of
    a table's first cross-polytope function are evenly distributed (+-3 % of the samples)."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    n, d, L = 4000, 100, 8
    data = util.uniform_sphere(n, d, 91)
    ix = cb.init_with_config(data, cb.Config(L, 1.0, 10, 0.9, "bits"))
    ix.set_clustering([0], np.zeros(n, np.uint64), [2.0])
    ix.set_option("seed", 77)
    ix.build()
    codes = ix.export(cl.X_TABLE_HASHES, 0, np.uint32).reshape(L, n)
    assert np.all(codes < (1 << 24))
    assert int(np.bitwise_or.reduce(codes.ravel())) == (1 << 24) - 1 and int(np.bitwise_and.reduce(codes.ravel())) == 0
    sk = ix.export(cl.X_SKETCHES, 0, np.uint64).reshape(n, 32)
    assert np.all(np.bitwise_or.reduce(sk, axis=0) == np.uint64(0xFFFFFFFFFFFFFFFF))
    assert np.all(np.bitwise_and.reduce(sk, axis=0) == 0)
    for t in range(L):
        first = np.bincount(codes[t] >> 16, minlength=256)   # the most significant of the three 8-bit function values
        assert np.all(np.abs(first - n / 256) <= 0.03 * n), t
    ix.close()


def _standalone_vs_oracle(oracle, n, d, L, k, delta, factor, seed, n_centers=5, nq=24):
    """Same comparison as smoke(): a stand-alone build on a shared function set handed in from the host, every query against the
    oracle over the same functions — ids, distance bits, candidates, distance computations, clusters visited."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    from oracle.pyoracle import Functions
    rng = np.random.default_rng(seed)
    data = util.planted(n, d, 400 + seed, n_centers=n_centers)
    queries = np.concatenate([util.planted_queries(data, nq - 4, 401 + seed), util.uniform_sphere(4, d, 402 + seed)]).astype(np.float32)
    sl = (d + 15) // 16 * 16
    m = int(np.ceil(np.log2(d)))
    fph = (24 + m) // (m + 1)
    planes = oracle.store_q15(rng.standard_normal((2048, d)).astype(np.float32), sl)
    signs = (rng.integers(0, 2, (L * fph, 3 << m)) * 2 - 1).astype(np.int8)
    index = cb.init_with_config(data, cb.Config(L, factor, k, delta, "edge"))
    index.set_functions(None, planes, signs, None)
    index.build()
    ids, dists, counts = index.search_batch(queries)
    ctr = index.counters(len(queries))
    K = index.num_clusters
    cen, asg = index.export(cl.X_CENTERS, 0, np.uint64), index.export(cl.X_ASSIGNMENT, 0, np.uint64)
    rad, brute = index.export(cl.X_RADII, 0, np.float32), index.export(cl.X_BRUTE, 0, np.uint8)
    assert not brute.all()
    oc_, oa_, _ = oracle.gmm(data, K)
    assert np.array_equal(cen, oc_) and np.array_equal(asg, oa_)
    fn = Functions(d, L, planes, signs, index.export(cl.X_EST, int(np.argmin(brute)), np.float32))
    oc = oracle.clann(data, k, delta, cen, asg, rad)
    for ci in range(K):
        if not brute[ci]:
            oc.build_cluster(ci, fn)
    for i in range(len(queries)):
        o_ids, o_d, _, o_ctr = oc.search(queries[i])
        c = int(counts[i])
        assert c == len(o_ids) and sorted(ids[i, :c].tolist()) == sorted(int(x) for x in o_ids), i
        assert np.array_equal(np.sort(dists[i, :c]), np.sort(o_d)), i
        assert int(ctr["candidates"][i]) == o_ctr["candidates"], i
        assert int(ctr["distance_computations"][i]) == o_ctr["distance_computations"], i
        assert int(ctr["clusters_visited"][i]) == o_ctr["visited"], i
    oc.free()
    index.close()


@pytest.mark.parametrize("d", [2, 3, 7, 300, 1000])
def test_unusual_row_widths(oracle, d):
    """The ends of the supported range. d > 256 (the documented limit is d <= 1024): storage rows wider than the register-resident
    query chunk, 512- / 1024-point FHT, the CUDA-core sketch kernel, no dense first-visit precompute. d = 2, 3, 7: 2- to 8-point
    FHT with 12 / 8 / 6 functions per table and code bits cut off the last one (d = 7: 6 x 4 bits; d = 3: 8 x 3), one 16-element
    storage row mostly padding, the 8-lanes-per-row k-center pass (d % 4 != 0)."""
    _standalone_vs_oracle(oracle, n=1500, d=d, L=8, k=10, delta=0.9, factor=0.2, seed=d)


@pytest.mark.parametrize("n,d,L,k,factor", [(6000, 32, 8, 1000, 0.04), (3000, 32, 8, 300, 0.05), (1200, 16, 1, 5, 0.1),
                                            (1200, 16, 300, 5, 0.1)])
def test_large_k_and_extreme_table_counts(oracle, n, d, L, k, factor):
    """k = 300 / 1000 inside PUFFINN clusters (1 024- / 2 048-slot MaxBuffer per warp, fewer warps per CTA), a single table, and
    300 tables (per-warp anchor / range arrays of 300 entries, ten-word stop masks)."""
    _standalone_vs_oracle(oracle, n=n, d=d, L=L, k=k, delta=0.9, factor=factor, seed=n + L + k, n_centers=3, nq=16)


def test_cluster_beyond_65536_rows(oracle):
    """One cluster of 70 000 rows (K = 1): local ids no longer fit 16 bits, so the opt-in u16 candidate streams and the dense
    first-visit memo must step aside (kernels_search.cu: nc > 65536 guards) while the default probe, whose table indices are 32-bit,
    answers as the oracle does — hundreds of thousands of candidates per query, long ranges at low depths."""
    _standalone_vs_oracle(oracle, n=70_000, d=16, L=4, k=10, delta=0.9, factor=0.005, seed=70_014, n_centers=1, nq=12)
