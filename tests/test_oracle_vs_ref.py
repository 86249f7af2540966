"""Pins the plain-C restatement (oracle/clann_oracle.c) against the REAL reference compiled from /root/reference
(oracle/_ref): stores, FHT, tables, anchors, ranges, result ids and the counters that depend on every candidate."""
import numpy as np
import pytest


def test_store_q15_all_small_dims(oracle, reflib):
    rng = np.random.default_rng(7)
    for d in [1, 2, 3, 5, 16, 17, 25, 31, 64, 96, 100, 101, 102, 103, 128, 130]:
        ix = reflib.index(d)
        for x in rng.standard_normal((40, d)).astype(np.float32) * rng.uniform(0.1, 5):
            assert np.array_equal(ix.store_q15(x), oracle.store_q15(x)), d
        assert not oracle.store_q15(np.zeros(d, np.float32)).any()  # zero vector stays zero (unit_vector.hpp:77)
        ix.free()


@pytest.mark.parametrize("m", [5, 7])
def test_fht_bit_exact_on_raw_floats(oracle, reflib, m):
    rng = np.random.default_rng(m)
    for _ in range(300):
        v = (rng.standard_normal(1 << m) * 10 ** rng.uniform(-3, 3)).astype(np.float32)
        assert np.array_equal(reflib.fht(v, m).view(np.uint32), oracle.fht(v, m).view(np.uint32))


@pytest.mark.parametrize("n,d,L", [(500, 25, 12), (700, 100, 20), (150, 3, 5)])
def test_index_build_and_search(oracle, reflib, n, d, L):
    rng = np.random.default_rng(n + d)
    X = rng.standard_normal((n, d)).astype(np.float32)
    ri = reflib.index(d, X, L, seed=100 + d)
    oi = oracle.index_import(ri.serialize())
    ob = oracle.index_build(oi.functions(), X)
    assert np.array_equal(ob.q15, oi.q15) and np.array_equal(ob.sketches, oi.sketches)
    assert np.array_equal(ob.hashes, oi.hashes) and np.array_equal(ob.indices, oi.indices)
    # sorted tables: (hash asc, id asc), 12 sentinels each side (prefixmap.hpp:215-226)
    h, i = ri.table(0)
    assert np.all(h[:12] == 0xFFFFFFFF) and np.all(h[-12:] == 0xFFFFFFFF)
    key = h[12:-12].astype(np.uint64) << 32 | i[12:-12]
    assert np.all(np.diff(key.astype(np.int64)) > 0)
    Q = rng.standard_normal((60, d)).astype(np.float32)
    for q in Q:
        q15 = oracle.store_q15(q)
        codes = oi.codes(q15)
        assert np.array_equal(ri.query_codes(q, L).astype(np.uint32), codes)
        assert np.array_equal(ri.query_sketches(q), oi.sketch(q15))
        ra, rr = ri.query_ranges(q, L)
        oa, orr = oi.query_ranges(codes)
        assert np.array_equal(ra, oa) and np.array_equal(rr, orr)
        for k, rec, ms in [(10, 0.9, float("-inf")), (1, 0.5, 0.6), (7, 0.95, 0.8)]:
            r_ids, rm = ri.search(q, k, rec, ms)
            o_ids, om = oi.search(q, k, rec, ms)
            assert np.array_equal(r_ids, o_ids)
            assert rm["distance_computations"] == om["distance_computations"] and rm["candidates"] == om["candidates"]
            maps = (24 - om["stop_depth"]) * L + om["stop_table"] if om["stop_depth"] else 0
            assert rm["hash_length"] == om["stop_depth"] and rm["considered_maps"] == maps
    ri.free(); oi.free(); ob.free()


@pytest.mark.parametrize("n,d,L", [(1500, 33, 24), (900, 100, 10), (60, 25, 4)])
def test_filter_type_none_and_simple(oracle, reflib, n, d, L):
    """Index::search with FilterType::None / FilterType::Simple (collection.hpp:22-34,671-765): ids in order, the depth the
    per-depth stop rule fired at (hash_length), considered_maps = (24 - depth + 1) L (:706-707,760-761), and counters that stay
    at zero (neither variant calls add_candidates / add_distance_computations). n = 60 takes the brute-force path (:550-555)."""
    rng = np.random.default_rng(n + L)
    centres = rng.standard_normal((5, d)).astype(np.float32)
    X = (centres[rng.integers(0, 5, n)] + 0.5 * rng.standard_normal((n, d))).astype(np.float32)
    ri = reflib.index(d, X, L, seed=300 + d)
    oi = oracle.index_import(ri.serialize())
    Q = (X[rng.integers(0, n, 40)] + 0.15 * rng.standard_normal((40, d))).astype(np.float32)
    depths = set()
    for ft in (1, 2):
        for k, rec in [(10, 0.9), (1, 0.5), (40, 0.95), (3, 0.2)]:
            for q in Q:
                r_ids, rm = ri.search(q, k, rec, filter_type=ft)
                o_ids, om = oi.search(q, k, rec, filter_type=ft)
                assert np.array_equal(r_ids, o_ids), (ft, k, rec)
                assert rm["hash_length"] == om["stop_depth"]
                assert rm["considered_maps"] == ((24 - om["stop_depth"] + 1) * L if om["stop_depth"] else 0)
                assert rm["distance_computations"] == 0 and rm["candidates"] == 0
                assert om["distance_computations"] == 0 and om["candidates"] == 0
                depths.add(om["stop_depth"])
    assert n < 100 or len(depths) > 2  # the cases stop at different depths
    ri.free(); oi.free()


def test_failure_probability_and_sketch_threshold(oracle, reflib):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((200, 100)).astype(np.float32)
    ri = reflib.index(100, X, 9, seed=5)
    oi = oracle.index_import(ri.serialize())
    for depth in (1, 7, 8, 9, 16, 23, 24):
        for t in (0, 1, 5, 9):
            for sim in (0.0, 0.3, 0.5, 0.5004, 0.77, 0.9999):
                mt = t if depth == 24 else 9
                a, b = ri.failure_probability(depth, t, mt, sim), oi.failure_probability(depth, t, mt, sim)
                assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
    for v in range(0, 65536, 97):
        sim = np.float32(v) / np.float32(65536)
        assert ri.max_sketch_diff(sim) == oracle.lib.orc_max_sketch_diff(float(sim))
    ri.free(); oi.free()


def test_small_index_brute_force_path(oracle, reflib):
    # fewer than 100 points: Index::search falls back to brute force (collection.hpp:550-555)
    rng = np.random.default_rng(2)
    X = rng.standard_normal((60, 25)).astype(np.float32)
    ri = reflib.index(25, X, 4, seed=9)
    oi = oracle.index_import(ri.serialize())
    for q in rng.standard_normal((20, 25)).astype(np.float32):
        assert np.array_equal(ri.search(q, 5, 0.9)[0], oi.search(q, 5, 0.9)[0])
    ri.free(); oi.free()


def test_clann_layer_restatements_agree(oracle, reflib):
    """gmm.rs / index.rs restated twice (C oracle, C++ binder over the real PUFFINN): same clustering, same results."""
    from tests import util
    data = util.planted(1500, 25, 3)
    K = oracle.num_clusters(0.4, len(data))
    c1, a1, r1 = oracle.gmm(data, K)
    c2, a2, r2 = reflib.gmm(data, K)
    assert np.array_equal(c1, c2) and np.array_equal(a1, a2) and np.array_equal(r1.view(np.uint32), r2.view(np.uint32))
    ref = reflib.clann(data, 8, 5, 0.9, c1, a1, r1, seed_base=77)
    ref.build_all()
    orc = oracle.clann(data, 5, 0.9, c1, a1, r1)
    for ci in range(K):
        s = ref.cluster_stream(ci)
        if s:
            orc.set_cluster_stream(ci, s)
    for q in util.planted_queries(data, 40, 4):
        ri, rd, ro, rc = ref.search(q)
        oi, od, oo, oc = orc.search(q)
        assert list(ri) == list(oi) and np.array_equal(rd, od) and np.array_equal(ro, oo) and rc == oc
        # per-visit rows (metrics/mod.rs:84-112): as many as visits, clusters in visiting order; the distance computations are
        # PUFFINN's counter (pinned above against the real PUFFINN) + one prune-test evaluation per visit after the first
        # + the list length of brute-force visits; a visit cannot add more points than the list it received (<= k)
        vi, vd, log = orc.search_visits(q, 32)
        assert list(vi) == list(oi) and np.array_equal(vd, od)
        assert len(log) == oc["visited"] and np.array_equal(log[:, 0], oo[: len(log)])
        brute_len = sum(min(5, int((a1 == c).sum())) for c in log[:, 0] if (a1 == c).sum() < 100)
        assert int(log[:, 2].sum()) == oc["distance_computations"] + (len(log) - 1) + brute_len
        assert np.all(log[:, 1] <= 5) and int(log[0, 1]) > 0
    ref.free(); orc.free()
