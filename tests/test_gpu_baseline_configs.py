"""Parity of the CUDA path with the oracle ON THE CONFIGURATIONS BASELINE.json NAMES, at their own sizes.

The scenarios of test_gpu_parity.py are small enough for the real reference to build every cluster; here the index has
the benchmarked shape (glove-100 / glove-25 shapes in full, d = 96 with k = 100 at three deltas) and the comparison is made
on a sample of queries, for which the oracle (oracle/clann_oracle.c, pinned against oracle/_ref) needs only the clusters
those queries visit. Per query the bar is the one of test_gpu_parity.py: identical visiting order, clusters visited,
`candidates`, `distance_computations`, ids up to 1e-5 ties and identical distance bits. On top of that the order-free traces
the C ABI exports are compared directly: anchors and all 24 ranges per (query, table) (prefixmap.hpp:36-57,267-304) and the
stop point (depth, table index) of the delta rule (collection.hpp:927-943) — BASELINE.md gate 5.

What is searched is exactly what bench.py times: stand-alone build, one function set shared by all clusters, the default
probe schedule (dense first-visit similarities + precomputed first-visit anchors), then the gather schedule as well.
"""
import os

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _functions(oracle, d, L, seed):
    """A function set drawn here so that the oracle and the device use the same one (SimHash planes in Q15, FHT signs)."""
    rng = np.random.default_rng(seed)
    sl = (d + 15) // 16 * 16
    m = int(np.ceil(np.log2(d)))
    fph = (24 + m) // (m + 1)
    planes = oracle.store_q15(rng.standard_normal((2048, d)).astype(np.float32), sl)
    signs = (rng.integers(0, 2, (L * fph, 3 << m)) * 2 - 1).astype(np.int8)
    return planes, signs


def _home_queries(data, assignment, clusters, per, seed, noise=0.05):
    """`per` planted queries (data point + noise) for each of the given home clusters."""
    rng = np.random.default_rng(seed)
    picks = []
    for c in clusters:
        members = np.flatnonzero(assignment == c)
        picks.append(rng.choice(members, size=min(per, len(members)), replace=False))
    picks = np.concatenate(picks)
    q = data[picks] + noise * rng.standard_normal((len(picks), data.shape[1])).astype(np.float32)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


def _between_queries(data, n, seed):
    """Midpoints of random pairs of points: queries that sit between clusters and walk several of them."""
    rng = np.random.default_rng(seed)
    a, b = rng.integers(0, len(data), n), rng.integers(0, len(data), n)
    q = data[a] + data[b]
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


def compare_with_oracle(ix, data, q, k, delta, oracle, independent_builds=2, trace_queries=24):
    """Search `q` on the device and replay every query in the oracle over the clusters it visited. Returns a dict of
    statistics; raises AssertionError on the first kind of difference."""
    from clann_b200 import _lib as cl
    nq = len(q)
    ix.set_option("visit_log", 16)   # the cluster granularity of the reference's metrics, recorded by the probe
    ids, dists, counts = ix.search_batch(q)
    ctr = ix.counters(nq)
    vlog = ix.visit_log(nq).copy()
    ix.set_option("visit_log", 0)
    K = ix.num_clusters
    L = ix.config.num_tables
    order = ix.export(cl.X_CLUSTER_ORDER, 0, np.uint32).reshape(nq, K)
    stop_points = ix.export(cl.X_STOP_POINTS, 0, np.uint32).reshape(nq, 2).copy()
    codes = None
    cen = ix.export(cl.X_CENTERS, 0, np.uint64)
    asg = ix.export(cl.X_ASSIGNMENT, 0, np.uint64)
    rad = ix.export(cl.X_RADII, 0, np.float32)
    brute = ix.export(cl.X_BRUTE, 0, np.uint8)
    vis = ctr["clusters_visited"]
    need = sorted({int(c) for i in range(nq) for c in order[i, : int(vis[i])] if not brute[c]})
    orc = oracle.clann(data, k, delta, cen, asg, rad)
    streams = {}
    for ci in need:
        streams[ci] = ix.export(cl.X_REFERENCE_STREAM, ci, np.uint8).tobytes()
        orc.set_cluster_stream(ci, streams[ci])
    # the exported streams are not taken on trust: the oracle rebuilds a few clusters from the raw rows with the same
    # functions (Q15 store, sketches, codes, stable sort) and must arrive at the same bytes
    for ci in need[:independent_builds]:
        oi = oracle.index_import(streams[ci])
        ob = oracle.index_build(oi.functions(), data[np.flatnonzero(asg == ci)])
        assert np.array_equal(ob.q15, oi.q15) and np.array_equal(ob.sketches, oi.sketches), f"cluster {ci}: rows / sketches differ"
        assert np.array_equal(ob.hashes, oi.hashes) and np.array_equal(ob.indices, oi.indices), f"cluster {ci}: tables differ"
        oi.free(); ob.free()
    bad = []
    for i in range(nq):
        o_ids, o_d, o_order, o_ctr = orc.search(q[i])
        # search_metrics_cluster rows (sqlite.rs:248-283): cluster, heap adds that returned true, distance computations incl. the
        # prune-test evaluation — per visit, in visiting order
        _, _, o_log = orc.search_visits(q[i], 16)
        g_log = vlog[i][: len(o_log)]
        assert np.array_equal(g_log[:, 0].astype(np.int64) - 1, o_log[:, 0].astype(np.int64)), (i, g_log, o_log)
        assert np.array_equal(g_log[:, 1:3].astype(np.uint64), o_log[:, 1:3]), (i, g_log, o_log)
        assert np.all(vlog[i][len(o_log):, 0] == 0) and np.all(g_log[:, 3] > 0)
        c = int(counts[i])
        g_ids, g_d = ids[i, :c], dists[i, :c]
        ok = (np.array_equal(order[i], o_order.astype(np.uint32))
              and int(vis[i]) == o_ctr["visited"]
              and int(ctr["candidates"][i]) == o_ctr["candidates"]
              and int(ctr["distance_computations"][i]) == o_ctr["distance_computations"]
              and util.same_ids_up_to_ties(g_ids, g_d, o_ids.astype(np.uint32), o_d)
              and np.array_equal(np.sort(g_d).view(np.uint32), np.sort(o_d).view(np.uint32)))
        if not ok:
            bad.append((i, int(vis[i]), o_ctr["visited"], int(ctr["candidates"][i]), o_ctr["candidates"],
                        int(ctr["distance_computations"][i]), o_ctr["distance_computations"]))
        assert np.all(ids[i, c:] == 0xFFFFFFFF) and np.all(np.isinf(dists[i, c:]))
    assert not bad, f"{len(bad)} of {nq} queries differ (i, vis gpu/orc, cand gpu/orc, dc gpu/orc): {bad[:8]}"
    # order-free traces (BASELINE.md gate 5): anchors, the 24 ranges per table, the stop point
    traced = 0
    by_home = {}
    for i in range(nq):
        home = int(order[i, 0])
        if not brute[home]:
            by_home.setdefault(home, []).append(i)
    for home, members in by_home.items():
        if traced >= trace_queries:
            break
        oi = oracle.index_import(streams[home])
        if codes is None:
            codes = {}
        qc = ix.export(cl.X_QUERY_CODES, home, np.uint32).reshape(nq, L)
        anchors = ix.export(cl.X_QUERY_ANCHORS, home, np.uint32).reshape(nq, L)
        ranges = ix.export(cl.X_QUERY_RANGES, home, np.uint32).reshape(nq, 24, L, 2)
        for i in members[:4]:
            a, r = oi.query_ranges(qc[i])
            assert np.array_equal(anchors[i], a), f"query {i}: anchors differ in cluster {home}"
            assert np.array_equal(ranges[i], r), f"query {i}: ranges differ in cluster {home}"
            if int(vis[i]) == 1:  # one PUFFINN visit: the stop point of the state is the one of this cluster
                _, info = oi.search(q[i], k, delta)
                assert (int(stop_points[i, 0]), int(stop_points[i, 1]) if stop_points[i, 0] else 0) == \
                       (info["stop_depth"], info["stop_table"] if info["stop_depth"] else 0), \
                    f"query {i}: stop point {tuple(stop_points[i])} vs oracle ({info['stop_depth']}, {info['stop_table']})"
            traced += 1
        oi.free()
    orc.free()
    return dict(clusters=len(need), visited_mean=float(np.mean(vis)), traced=traced,
                candidates_mean=float(np.mean(ctr["candidates"])), dc_mean=float(np.mean(ctr["distance_computations"])))


def _standalone_index(oracle, data, L, k, delta, fn_seed, name):
    import clann_b200 as cb
    planes, signs = _functions(oracle, data.shape[1], L, fn_seed)
    ix = cb.init_with_config(data, cb.Config(L, 0.4, k, delta, name))
    ix.set_functions(None, planes, signs, None)  # shared set; collision estimates from the device Monte-Carlo
    ix.build()
    return ix


def _both_schedules(ix, data, q, k, delta, oracle, **kw):
    """The default schedule (dense first-visit similarities + first-visit anchors) and the two alternatives."""
    from clann_b200 import _lib as cl
    stats = compare_with_oracle(ix, data, q, k, delta, oracle, **kw)
    # the gather schedule (no precompute at all), and the opt-in first-visit candidate stream
    for knob, value, restore in (("dense_sims", 0, 1), ("first_stream", 1, 0)):
        cl.tune(knob, value)
        try:
            other = compare_with_oracle(ix, data, q, k, delta, oracle, independent_builds=0, trace_queries=0)
        finally:
            cl.tune(knob, restore)
        assert other["candidates_mean"] == stats["candidates_mean"] and other["dc_mean"] == stats["dc_mean"], knob
    return stats


def test_glove100_full_shape_against_oracle(oracle):
    """BASELINE.json configs[2] (the bench line): 1 183 514 x 100, L = 84, K = 435, k = 10, delta = 0.9."""
    from clann_b200 import _lib as cl
    n, d = 1_183_514, 100
    data = util.planted(n, d, 42)
    ix = _standalone_index(oracle, data, 84, 10, 0.9, 1001, "glove-100-shape")
    assert ix.num_clusters == 435
    asg = ix.export(cl.X_ASSIGNMENT, 0, np.uint64)
    homes = np.random.default_rng(7).choice(435, 24, replace=False)
    q = np.concatenate([_home_queries(data, asg, homes, 12, 43), _between_queries(data, 12, 44)])
    stats = _both_schedules(ix, data, q, 10, 0.9, oracle)
    assert stats["traced"] >= 20 and stats["visited_mean"] >= 1.0
    ix.close()


def test_glove25_full_shape_against_oracle(oracle):
    """BASELINE.json configs[1]: 1 183 514 x 25 (m = 5, four functions per code). The planted blobs overlap at d = 25, a
    query walks ~20 clusters, and the adaptive schedule moves to the gather path after the first batch: both are compared."""
    from clann_b200 import _lib as cl
    n, d = 1_183_514, 25
    data = util.planted(n, d, 42)
    ix = _standalone_index(oracle, data, 84, 10, 0.9, 1002, "glove-25-shape")
    asg = ix.export(cl.X_ASSIGNMENT, 0, np.uint64)
    homes = np.random.default_rng(8).choice(ix.num_clusters, 3, replace=False)
    q = _home_queries(data, asg, homes, 10, 45)
    stats = _both_schedules(ix, data, q, 10, 0.9, oracle, trace_queries=8)
    assert stats["visited_mean"] > 3.0, stats   # the shape that forces the adaptive gather schedule
    # third call: the adaptive choice (made from the statistics of the batches above) must not change anything either
    compare_with_oracle(ix, data, q[:10], 10, 0.9, oracle, independent_builds=0, trace_queries=0)
    ix.close()


@pytest.mark.parametrize("delta", [0.8, 0.9, 0.95])
def test_deep96_k100_deltas_against_oracle(oracle, delta):
    """BASELINE.json configs[3]/[4] arithmetic (d = 96: SL = 96, no padding; k = 100: 256-slot MaxBuffer) at the three deltas of
    the sweep, on a 300 000-point sample of the shape."""
    from clann_b200 import _lib as cl
    data = util.planted(300_000, 96, 46)
    ix = _standalone_index(oracle, data, 84, 100, delta, 1003, "deep-96-shape")
    asg = ix.export(cl.X_ASSIGNMENT, 0, np.uint64)
    homes = np.random.default_rng(9).choice(ix.num_clusters, 8, replace=False)
    q = np.concatenate([_home_queries(data, asg, homes, 10, 47), _between_queries(data, 8, 48)])
    stats = _both_schedules(ix, data, q, 100, delta, oracle, independent_builds=1, trace_queries=8)
    assert stats["traced"] >= 4
    ix.close()
