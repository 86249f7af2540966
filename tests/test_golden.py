"""The oracle against the committed golden fixtures (tests/golden/*.npz, generated from the real reference by
tests/golden/make_golden.py). These run everywhere, including the GPU box where /root/reference does not exist."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["puffinn_d25", "puffinn_d100"]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden_build(oracle, name):
    g = load(name)
    oi = oracle.index_import(g["stream"].tobytes())
    ob = oracle.index_build(oi.functions(), g["data"])
    assert np.array_equal(ob.q15, oi.q15)
    assert np.array_equal(ob.sketches, oi.sketches)
    assert np.array_equal(ob.hashes, oi.hashes) and np.array_equal(ob.indices, oi.indices)
    oi.free(); ob.free()


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden_queries(oracle, name):
    g = load(name)
    L = int(g["L"])
    oi = oracle.index_import(g["stream"].tobytes())
    q15 = oracle.store_q15(g["queries"])
    assert np.array_equal(q15, g["query_q15"])
    for qi, q in enumerate(g["queries"]):
        codes = oi.codes(q15[qi])
        assert np.array_equal(codes, g["query_codes"][qi])
        assert np.array_equal(oi.sketch(q15[qi]), g["query_sketches"][qi])
        a, r = oi.query_ranges(codes)
        assert np.array_equal(a, g["anchors"][qi]) and np.array_equal(r, g["ranges"][qi])
        for si, (k, rec, ms) in enumerate(g["searches"]):
            ids, m = oi.search(q, int(k), float(rec), float(ms))
            cnt = int(g["res_cnt"][si, qi])
            assert np.array_equal(ids, g["res_ids"][si, qi, :cnt])
            dc, cand, hl, maps = (int(x) for x in g["res_met"][si, qi])
            assert m["distance_computations"] == dc and m["candidates"] == cand and m["stop_depth"] == hl
            assert ((24 - hl) * L + m["stop_table"] if hl else 0) == maps
    oi.free()


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden_filter_types(oracle, name):
    """FilterType::None / FilterType::Simple (collection.hpp:671-765) on the same stored index: ids in the reference's order and
    the depth its per-depth stop rule fired at (tests/golden/make_golden.py filters)."""
    g, f = load(name), load(name + "_filters")
    oi = oracle.index_import(g["stream"].tobytes())
    for fi, ft in enumerate(f["filter_types"]):
        for si, (k, rec) in enumerate(f["searches"]):
            for qi, q in enumerate(g["queries"]):
                ids, m = oi.search(q, int(k), float(rec), filter_type=int(ft))
                cnt = int(f["res_cnt"][fi, si, qi])
                assert np.array_equal(ids, f["res_ids"][fi, si, qi, :cnt]), (ft, si, qi)
                assert m["stop_depth"] == int(f["res_depth"][fi, si, qi])
    oi.free()
