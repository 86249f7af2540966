"""The oracle against every known-answer test the reference holds for this path:
format_test.hpp:10-33, similarity_measure_test.hpp:13-41, maxbuffer_test.hpp:10-73, dataset_test.hpp:13-30,
math_test.hpp:13-30, hash_test.hpp:140-151, filterer_test.hpp:12-42, src/core/heap.rs:55-161, src/core/index.rs:695-748.
"""
import numpy as np
import pytest


def test_to_16bit_fixed_point(oracle):  # format_test.hpp:10-17
    f = oracle.lib.orc_to_q15
    assert f(0.99999) == 32767 and f(1.0) == 32767 and f(-1.0) == -32768 and f(0.0) == 0
    assert f(0.5) == 0x4000 and f(-0.5) == np.int16(np.uint16(0xC000))


def test_from_16bit_fixed_point(oracle):  # format_test.hpp:19-26
    f = oracle.lib.orc_from_q15
    assert f(0) == 0.0 and f(0x4000) == 0.5
    assert f(int(np.int16(np.uint16(0xA000)))) == -0.75 and f(int(np.int16(np.uint16(0x8000)))) == -1.0
    assert f(0x7FFF) == np.float32(32767) / np.float32(32768)


def test_pad_dimensions(oracle):  # format_test.hpp:28-33
    assert [oracle.lib.orc_storage_len(d) for d in (0, 1, 16, 17)] == [0, 16, 16, 32]


def test_compute_similarity_known_answer(oracle):  # similarity_measure_test.hpp:13-41
    a = oracle.store_q15(np.array([0.2, 0.4, 0.0, 0.8, 0.4], np.float32))
    b = oracle.store_q15(np.array([0.4, 0.0, 0.8, 0.2, 0.4], np.float32))
    assert abs(oracle.similarity(a, b) - 0.7) <= 1e-4
    l1, l2 = np.zeros(64, np.float32), np.zeros(64, np.float32)
    l1[[2, 16, 27, 63]] = [0.2, 0.4, 0.8, 0.4]
    l2[[2, 27, 51, 63]] = [0.4, 0.2, 0.8, 0.4]
    assert abs(oracle.similarity(oracle.store_q15(l1), oracle.store_q15(l2)) - 0.7) <= 1e-4


MAXBUFFER_CASES = [  # maxbuffer_test.hpp:10-73 (k, inserts, best_entries, smallest_value or None)
    (0, [(1, 0.5)], [], None),
    (2, [], [], None),
    (2, [(2, 0.6)], [(2, 0.6)], 0.0),
    (2, [(100, 0.1)], [(100, 0.1)], 0.0),
    (2, [(100, 0.1), (50, 0.2), (105, 0.3)], [(105, 0.3), (50, 0.2)], None),
    (2, [(1, 0.1), (2, 0.5), (3, 0.05), (4, 0.07), (5, 0.5), (6, 0.9), (7, 0.7), (8, 0.8)], [(6, 0.9), (8, 0.8)], None),
    (2, [(1, 0.1), (1, 0.1), (1, 0.1)], [(1, 0.1)], None),
    (2, [(1, -5.0), (2, 1.2)], [(2, 1.0)], None),
]


@pytest.mark.parametrize("k,inserts,expected,minval", MAXBUFFER_CASES)
def test_maxbuffer_known_answers(oracle, k, inserts, expected, minval):
    ids = [i for i, _ in inserts]
    vals = [v for _, v in inserts]
    oi, ov, mv = oracle.maxbuffer_run(k, ids, vals)
    assert list(oi) == [i for i, _ in expected]
    assert np.allclose(ov, [v for _, v in expected])
    if minval is not None:
        assert mv == minval


def test_maxbuffer_minval_after_filter(oracle):
    # maxbuffer_test.hpp:37-45,47-60: smallest_value() is read after best_entries() there, i.e. after a filter
    _, ov, _ = oracle.maxbuffer_run(2, [100, 50, 105], [0.1, 0.2, 0.3])
    assert ov[-1] == np.float32(0.2)
    _, ov, _ = oracle.maxbuffer_run(2, [1, 2, 3, 4, 5, 6, 7, 8], [0.1, 0.5, 0.05, 0.07, 0.5, 0.9, 0.7, 0.8])
    assert ov[-1] == np.float32(0.8)


def test_dataset_accessor(oracle):  # dataset_test.hpp:13-30
    assert oracle.lib.orc_storage_len(3) == 16
    rows = oracle.store_q15(np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32))
    assert rows.shape == (3, 16)
    assert rows[0, 0] == 32767 and rows[1, 2] == -32768 and rows[2, 1] == 32767
    assert not rows[:, 3:].any()


def test_dot_simple_equals_avx2(oracle, reflib):  # math_test.hpp:13-30
    rng = np.random.default_rng(0)
    for _ in range(100):
        a = oracle.store_q15(rng.standard_normal(100).astype(np.float32))
        b = oracle.store_q15(rng.standard_normal(100).astype(np.float32))
        assert reflib.dot_i16(a, b) == reflib.dot_i16(a, b, simple=True) == oracle.dot_i16(a, b)


def test_bits_per_function(oracle):  # hash_test.hpp:140-151
    assert oracle.lib.orc_ceil_log(100) + 1 == 8
    assert oracle.lib.orc_ceil_log(25) + 1 == 6 and oracle.lib.orc_ceil_log(128) + 1 == 8


def test_sketch_filter_semantics(oracle):  # filterer_test.hpp:12-42
    # equal vectors have sketch distance 0 <= max_sketch_diff(1.0); opposite vectors differ in every bit
    assert oracle.lib.orc_max_sketch_diff(1.0) == 0
    assert oracle.lib.orc_max_sketch_diff(0.0) == 64
    assert oracle.lib.orc_max_sketch_diff(0.5) == 32


def test_heap_rs_unit_tests(oracle):  # src/core/heap.rs:55-161
    lst, added, _ = oracle.topk_run(3, [2.5, 1.5], [0, 1])
    assert lst == [(1.5, 1), (2.5, 0)] and added.all()
    lst, _, _ = oracle.topk_run(3, [3.0, 2.0, 1.0, 0.5], [0, 1, 2, 3])
    assert len(lst) == 3 and (1.0, 2) in lst and (2.0, 1) in lst and (0.5, 3) in lst
    lst, added, _ = oracle.topk_run(3, [3.0, 2.0, 1.0, 4.0], [0, 1, 2, 3])
    assert not added[3] and len(lst) == 3 and (4.0, 3) not in lst
    _, _, top = oracle.topk_run(2, [2.0, 1.0], [1, 2])
    assert top == (1, 2.0)
    _, _, top = oracle.topk_run(2, [2.0, 1.0, 0.5], [1, 2, 3])
    assert top == (2, 1.0)
    lst, _, top = oracle.topk_run(3, [], [])
    assert lst == [] and top is None


def test_sort_cluster_known_answer(oracle):  # src/core/index.rs:695-748 — the only known-answer test of the CLANN layer
    pts = np.array([
        [0.1, 0.9, 0.4], [0.7, 0.2, 0.6], [0.5, 0.3, 0.9], [0.8, 0.4, 0.1], [0.2, 0.1, 0.8], [0.9, 0.8, 0.3], [0.3, 0.6, 0.5],
        [0.4, 0.3, 0.7], [0.1, 0.2, 0.9], [0.6, 0.7, 0.8], [0.2, 0.8, 0.1], [0.9, 0.2, 0.4], [0.3, 0.5, 0.6], [0.1, 0.9, 0.2],
        [0.7, 0.4, 0.6], [0.8, 0.3, 0.2], [0.4, 0.6, 0.3], [0.2, 0.7, 0.9], [0.9, 0.4, 0.8], [0.5, 0.1, 0.3]], np.float32)
    order = oracle.sort_clusters(pts, [6, 3, 17], [0.1, 0.0, 0.7])
    assert list(order) == [2, 0, 1]


def test_num_clusters_rule(oracle):  # src/core/index.rs:78-80
    assert oracle.num_clusters(0.4, 10_000) == 40
    assert oracle.num_clusters(0.4, 1_183_514) == 435
    assert oracle.num_clusters(0.4, 10_000_000) == 1264
    assert oracle.num_clusters(0.4, 100_000_000) == 4000
    assert oracle.num_clusters(0.0001, 10) == 1


def _unrolled_dot_f32(x, y):
    """ndarray 0.16.1 numeric_util::unrolled_dot written a second time, step by step in numpy float32 (every product and every
    sum rounded to f32, no FMA): eight partial sums over chunks of 8, (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then the tail in order."""
    f = np.float32
    p = [f(0)] * 8
    i = 0
    while i + 8 <= len(x):
        for j in range(8):
            p[j] = f(p[j] + f(x[i + j] * y[i + j]))
        i += 8
    s = f(0)
    for a, b in ((0, 4), (1, 5), (2, 6), (3, 7)):
        s = f(s + f(p[a] + p[b]))
    for j in range(i, len(x)):
        s = f(s + f(x[j] * y[j]))
    return s


def test_ndarray_dot_and_distance_restatements_agree(oracle):
    """The L3 arithmetic (angulardata.rs:12-43 over ndarray's dot) cannot be pinned against Rust here; this pins the C restatement
    against an independent one of the published algorithm: bit-equal dots for every length 1..40 and the benchmarked dimensions,
    and bit-equal distances (norm = sqrt(dot(x, x)), query norm = sequential sum of squares, 1 - dot / (|x| |q|))."""
    rng = np.random.default_rng(77)
    for d in list(range(1, 41)) + [96, 100, 128, 200]:
        for _ in range(6):
            x = (rng.standard_normal(d) * 10 ** rng.uniform(-2, 2)).astype(np.float32)
            y = rng.standard_normal(d).astype(np.float32)
            got = np.float32(oracle.lib.orc_ndarray_dot(x.ctypes.data, y.ctypes.data, d))
            assert got.view(np.uint32) == _unrolled_dot_f32(x, y).view(np.uint32), d
            f = np.float32
            xn = f(np.sqrt(_unrolled_dot_f32(x, x)))
            s = f(0)
            for v in y:
                s = f(s + f(v * v))
            want = f(f(1) - f(_unrolled_dot_f32(x, y) / f(xn * f(np.sqrt(s)))))
            assert np.float32(oracle.distance_point(x, y)).view(np.uint32) == want.view(np.uint32), d


def test_gmm_restatements_agree(oracle):
    """greedy_minimum_maximum (gmm.rs:21-62) over AngularData::distance (angulardata.rs:12-27) written a second time in numpy
    float32 on top of the step-by-step dot above: first-max arg-max (gmm.rs:5-15), strict-< reassignment, radii by f32::max.
    Small cases incl. duplicated rows (ties in the arg-max and in the reassignment) and n <= k."""
    f = np.float32
    rng = np.random.default_rng(5)
    for n, d, K in [(120, 12, 7), (90, 25, 11), (64, 9, 5)]:
        data = rng.standard_normal((n, d)).astype(np.float32)
        data[n // 2] = data[3]          # duplicates: equal distances everywhere
        data[n // 2 + 1] = data[3]
        norms = [f(np.sqrt(_unrolled_dot_f32(r, r))) for r in data]

        def dist_to(j):
            return [f(f(1) - f(_unrolled_dot_f32(data[i], data[j]) / f(norms[i] * norms[j]))) for i in range(n)]

        centers, assign = [0], [0] * n
        dist = dist_to(0)
        for idx in range(1, K):
            far, m = 0, dist[0]
            for i in range(1, n):
                if dist[i] > m:
                    far, m = i, dist[i]
            centers.append(far)
            nd = dist_to(far)
            for i in range(n):
                if nd[i] < dist[i]:
                    assign[i], dist[i] = idx, nd[i]
        radii = [f(0)] * K
        for i in range(n):
            radii[assign[i]] = max(radii[assign[i]], dist[i])
        c, a, r = oracle.gmm(data, K)
        assert list(c) == centers and list(a) == assign
        assert np.array_equal(np.asarray(radii, np.float32).view(np.uint32), r.view(np.uint32))
    c, a, r = oracle.gmm(rng.standard_normal((4, 6)).astype(np.float32), 9)   # gmm.rs:26-31
    assert list(c) == [0, 1, 2, 3] and list(a) == [0, 1, 2, 3] and not r.any()


def _golden_index(oracle, name="puffinn_d100"):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    return oracle.index_import(g["stream"].tobytes())


def test_independent_hashes_use_every_bit(oracle):
    """hash_source_test.hpp:13-45,76-91 ("Independent hashes", FHTCrossPolytopeHash, 24 bits): over random vectors every
    concatenated hash is below 2^24 and every one of the 24 bits is set at least once — on the reference-drawn function set of
    the golden index (d = 100: 3 functions of 8 bits per table)."""
    oi = _golden_index(oracle)
    rng = np.random.default_rng(3)
    seen = 0
    for v in rng.standard_normal((100, 100)).astype(np.float32):
        codes = oi.codes(oracle.store_q15(v))
        assert np.all(codes < (1 << 24))
        seen |= int(np.bitwise_or.reduce(codes))
    assert seen == (1 << 24) - 1
    oi.free()


def test_all_filtering_bits_used(oracle):
    """filterer_test.hpp:44-70: the sketches of 20 random vectors set every one of the 64 bit positions at least once (and,
    stronger, every position of every one of the 32 sketches takes both values)."""
    oi = _golden_index(oracle)
    rng = np.random.default_rng(4)
    sk = np.stack([oi.sketch(oracle.store_q15(v)) for v in rng.standard_normal((20, 100)).astype(np.float32)])
    assert int(np.bitwise_or.reduce(sk.ravel())) == 0xFFFFFFFFFFFFFFFF
    assert np.all(np.bitwise_or.reduce(sk, axis=0) == np.uint64(0xFFFFFFFFFFFFFFFF))
    assert np.all(np.bitwise_and.reduce(sk, axis=0) == 0)
    # filterer_test.hpp:12-42: a vector passes against itself at max_sketch_diff(1.0) = 0, its negation differs in every bit
    v = rng.standard_normal(100).astype(np.float32)
    a, b = oi.sketch(oracle.store_q15(v)), oi.sketch(oracle.store_q15(-v))
    differing = sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(a, b))
    assert differing >= 32 * 64 - 32   # a plane whose rounded dot is exactly 0 gives the same bit (>= 0) to v and -v
    oi.free()


def test_fht_cross_polytope_collision_probability(oracle):
    """hash_test.hpp:62-97,128-130 ("FHTCrossPolytope collision probability", d = 100): over 10 000 pairs of random unit vectors
    the number of collisions of one 8-bit function stays within 2 % of the sum of the tabulated probabilities
    (crosspolytope.hpp:116-118: estimates[bits][sim / eps]) — the table the stop rule is built on. Also "evenly distributed"
    (:40-60) in its 8-bit form: no outcome of a function takes more than 3 % of the samples beyond its share."""
    oi = _golden_index(oracle)
    fn = oi.functions()
    cfn = fn.c_struct()
    import ctypes as C
    rng = np.random.default_rng(6)
    n_fn = fn.L * fn.fph
    N = 10_000
    A = oracle.store_q15(rng.standard_normal((N, 100)).astype(np.float32))
    B = oracle.store_q15(rng.standard_normal((N, 100)).astype(np.float32))
    prob_sum, actual = 0.0, 0
    counts = np.zeros(1 << fn.bpf, np.int64)
    for i in range(N):
        f = i % n_fn
        ha = oracle.lib.orc_fht_hash(C.byref(cfn), f, A[i].ctypes.data)
        hb = oracle.lib.orc_fht_hash(C.byref(cfn), f, B[i].ctypes.data)
        assert ha < (1 << fn.bpf) and hb < (1 << fn.bpf)
        counts[ha] += 1
        sim = np.float32(oracle.similarity(A[i], B[i]))
        prob_sum += float(fn.est[fn.bpf][int(np.float32(sim / fn.eps))])
        actual += int(ha == hb)
    assert abs(prob_sum - actual) <= 0.02 * N, (prob_sum, actual)
    assert np.all(np.abs(counts - N / counts.size) <= 0.03 * N)
    # the same bound where collisions are frequent (random pairs collide 0.4 % of the time): correlated pairs, cosine ~0.6-0.95
    M = 4000
    X = rng.standard_normal((M, 100)).astype(np.float32)
    Y = X + rng.uniform(0.3, 1.2, (M, 1)).astype(np.float32) * rng.standard_normal((M, 100)).astype(np.float32)
    X15, Y15 = oracle.store_q15(X), oracle.store_q15(Y)
    prob_sum, actual = 0.0, 0
    for i in range(M):
        f = i % n_fn
        sim = np.float32(oracle.similarity(X15[i], Y15[i]))
        prob_sum += float(fn.est[fn.bpf][int(np.float32(sim / fn.eps))])
        actual += int(oracle.lib.orc_fht_hash(C.byref(cfn), f, X15[i].ctypes.data) == oracle.lib.orc_fht_hash(C.byref(cfn), f, Y15[i].ctypes.data))
    assert actual > 0.05 * M and abs(prob_sum - actual) <= 0.02 * M, (prob_sum, actual)
    oi.free()
