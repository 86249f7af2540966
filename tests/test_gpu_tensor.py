"""The tensor-pipe kernels (kernels_tc.cu: tcgen05.mma into TMEM, TMA operand loads) against the CUDA-core kernels they
replace and against the oracle: results must be bit-identical, including on inputs built to hit the exact-evaluation band and
the int16 wrap-around of the reference's dot product (math.hpp:37-44)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _adversarial(oracle, d, n, seed):
    """Planted rows plus: rows equal to hyperplanes (dot = +-32767-ish: the wrap band), their negations, near-orthogonal rows
    (the undecided band), axis-aligned rows (a component of 32767), a zero row, duplicates."""
    rng = np.random.default_rng(seed)
    raw_planes = rng.standard_normal((2048, d)).astype(np.float32)
    sl = (d + 15) // 16 * 16
    planes = oracle.store_q15(raw_planes, sl)
    data = util.planted(n, d, seed + 1, n_centers=3)
    unit = raw_planes / np.linalg.norm(raw_planes, axis=1, keepdims=True)
    data[0:40] = unit[0:40]
    data[40:60] = -unit[100:120]
    a, b = unit[200:230], unit[300:330]
    orth = b - (np.sum(a * b, axis=1, keepdims=True)) * a          # orthogonal to plane 200 + i
    data[60:90] = orth / np.linalg.norm(orth, axis=1, keepdims=True)
    eye = np.zeros((min(d, 16), d), np.float32)
    eye[np.arange(min(d, 16)), np.arange(min(d, 16))] = 1.0
    data[90:90 + len(eye)] = eye
    data[110] = 0.0
    data[111] = data[112]
    return data.astype(np.float32), planes


@pytest.mark.parametrize("d", [25, 96, 100, 128, 200])
def test_tensor_sketches_equal_cuda_core_and_oracle(oracle, d):
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    L = 6
    n = 3000
    data, planes = _adversarial(oracle, d, n, 100 + d)
    m = int(np.ceil(np.log2(d)))
    fph = (24 + m) // (m + 1)
    signs = (np.random.default_rng(5).integers(0, 2, (L * fph, 3 << m)) * 2 - 1).astype(np.int8)
    queries = np.concatenate([data[:130], util.uniform_sphere(70, d, 7) * 2.5]).astype(np.float32)
    got = {}
    for tc in (1, 0):
        cl.tune("tc_sketch", tc)
        try:
            ix = cb.init_with_config(data, cb.Config(L, 0.05, 5, 0.9, "tensor"))
            ix.set_functions(None, planes, signs, None)
            ix.build()
            K = ix.num_clusters
            brute = ix.export(cl.X_BRUTE, 0, np.uint8)
            sk = {int(c): ix.export(cl.X_SKETCHES, int(c), np.uint64).copy() for c in range(K) if not brute[c]}
            ix.search_batch(queries)
            qs = ix.export(cl.X_QUERY_SKETCHES, int(np.argmin(brute)), np.uint64).reshape(len(queries), 32).copy()
            got[tc] = (sk, qs)
            ix.close()
        finally:
            cl.tune("tc_sketch", 1)
    assert got[1][0].keys() == got[0][0].keys() and len(got[1][0]) > 0
    for c in got[1][0]:
        assert np.array_equal(got[1][0][c], got[0][0][c]), f"cluster {c}: tensor-pipe sketches differ from the CUDA-core kernel"
    assert np.array_equal(got[1][1], got[0][1])
    # and against the oracle (filterer.hpp:76-102 restated), query side
    from oracle.pyoracle import Functions
    est = np.ones((m + 2, 201), np.float32)
    fn = Functions(d, L, planes, signs, est)
    oi = oracle.index_build(fn, data[:200])
    q15 = oracle.store_q15(queries)
    for i in range(len(queries)):
        assert np.array_equal(got[1][1][i], oi.sketch(q15[i])), f"query {i}"
    oi.free()


@pytest.mark.parametrize("shape", [(40_000, 100, 10), (25_000, 64, 10), (20_000, 200, 5)])
def test_tensor_centre_scoring_changes_nothing(shape):
    """Query x centre scoring as a tf32 tensor-core GEMM screen + exact evaluation of the 32 nearest candidates
    (k_center_gemm_tc, k_center_refine; index.rs:592-600) against the all-exact CUDA-core kernel: identical visiting order,
    ids, distance bits and counters — for planted queries (one or two clusters), for uniform ones that walk far beyond the 32
    exact candidates (the probe re-evaluates their row), and for a zero query (NaN distances)."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    n, d, k = shape
    data = util.planted(n, d, 300 + d)
    q = np.concatenate([util.planted_queries(data, 300, 301), util.uniform_sphere(60, d, 302), np.zeros((1, d), np.float32)])
    ix = cb.init_with_config(data, cb.Config(20, 0.4, k, 0.9, "centre"))
    ix.set_option("seed", 3)
    ix.build()
    K = ix.num_clusters
    assert K > 40   # more centres than exact candidates per query
    res = {}
    for tc in (1, 0):
        cl.tune("tc_center", tc)
        cl.tune("dense_adaptive", 0)   # keep the screen on whatever the previous batch did
        try:
            ids, dists, counts = ix.search_batch(q)
            ctr = ix.counters(len(q))
            order = ix.export(cl.X_CLUSTER_ORDER, 0, np.uint32).reshape(len(q), K).copy()
            res[tc] = (ids, dists.view(np.uint32), counts, ctr["candidates"], ctr["distance_computations"], ctr["clusters_visited"], order)
        finally:
            cl.tune("tc_center", 1)
            cl.tune("dense_adaptive", 1)
    assert res[0][5].max() > 32    # some query does walk past the exact candidates
    for a, b in zip(res[1], res[0]):
        assert np.array_equal(a, b)
    ix.close()
