"""puffinn::Index::search with FilterType::None / FilterType::Simple (collection.hpp:22-34,671-765) on the device, through
clann_puffinn_search: ids in the reference's order and the depth at which the per-depth stop rule fired, against the golden
fixtures, the real reference (oracle/_ref) and the oracle."""
import os

import numpy as np
import pytest

from tests import util
from tests.test_gpu_more import _write_record_file

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load_stream(cb, tmp_path, stream: bytes, tag: str):
    path = str(tmp_path / f"{tag}.clb2")
    _write_record_file(path, "index_0", stream)
    return cb.PuffinnIndex.new_from_file(path, 0)


@pytest.mark.parametrize("name", ["puffinn_d25", "puffinn_d100"])
def test_filter_types_match_golden_fixture(name, tmp_path):
    """The reference's own answers (tests/golden/make_golden.py filters) on the stored index."""
    import clann_b200 as cb
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    f = np.load(os.path.join(GOLDEN, name + "_filters.npz"))
    index = _load_stream(cb, tmp_path, g["stream"].tobytes(), name)
    for fi, ft in enumerate(f["filter_types"]):
        for si, (k, rec) in enumerate(f["searches"]):
            for qi, q in enumerate(g["queries"]):
                ids, depth = index.search_filtered(q, int(k), float(rec), int(ft))
                cnt = int(f["res_cnt"][fi, si, qi])
                assert ids == f["res_ids"][fi, si, qi, :cnt].tolist(), (int(ft), si, qi)
                assert depth == int(f["res_depth"][fi, si, qi]), (int(ft), si, qi)
                assert cb.api.get_distance_computations() == 0  # neither variant counts (only collection.hpp:865,904,921 do)
    # filter_type 0 is the default path: the call CPUFFINN_search_cosine makes
    for qi, q in enumerate(g["queries"][:10]):
        ids, depth = index.search_filtered(q, 10, 0.9, 0)
        cnt = int(g["res_cnt"][0, qi])
        assert sorted(ids) == sorted(g["res_ids"][0, qi, :cnt].tolist())
        assert depth == int(g["res_met"][0, qi, 2])
        assert cb.api.get_distance_computations() == int(g["res_met"][0, qi, 0])


@pytest.mark.parametrize("n,d,L", [(3000, 64, 12), (1500, 33, 40), (60, 25, 4)])
def test_filter_types_match_reference_and_oracle(oracle, reflib, tmp_path, n, d, L):
    """An index built by the real reference, loaded on the device from its serialized stream: every (filter type, k, recall)
    returns the reference's ids in its order and stops at its depth; k = 100 exercises the 256-slot MaxBuffer, n = 60 the
    brute-force path (collection.hpp:550-555)."""
    import clann_b200 as cb
    rng = np.random.default_rng(n + L)
    centres = rng.standard_normal((6, d)).astype(np.float32)
    data = (centres[rng.integers(0, 6, n)] + 0.5 * rng.standard_normal((n, d))).astype(np.float32)
    ri = reflib.index(d, data, L, seed=900 + d)
    stream = ri.serialize()
    oi = oracle.index_import(stream)
    index = _load_stream(cb, tmp_path, stream, f"ref{n}")
    queries = (data[rng.integers(0, n, 48)] + 0.1 * rng.standard_normal((48, d))).astype(np.float32)
    depths = set()
    for ft in (1, 2):
        for k, rec in [(10, 0.9), (100, 0.95), (1, 0.5), (5, 0.2)]:
            for qi, q in enumerate(queries):
                ids, depth = index.search_filtered(q, k, rec, ft)
                r_ids, rm = ri.search(q, k, rec, filter_type=ft)
                o_ids, om = oi.search(q, k, rec, filter_type=ft)
                assert ids == r_ids.tolist() == o_ids.tolist(), (ft, k, rec, qi)
                assert depth == rm["hash_length"] == om["stop_depth"], (ft, k, rec, qi)
                depths.add(depth)
    assert n < 100 or len(depths) > 2
    ri.free(); oi.free()


def test_filter_types_on_device_built_index(oracle, tmp_path):
    """Functions drawn and tables built on the device (CPUFFINN_index_create / insert / rebuild); the saved record is the
    reference's serialization, so the oracle replays the same index. Also: argument errors never fall through."""
    import clann_b200 as cb
    from clann_b200 import _lib as cl
    data = util.planted(2500, 48, 61, n_centers=4)
    index, _ = cb.PuffinnIndex.new(cb.AngularData(data), 16)
    path = str(tmp_path / "dev.clb2")
    index.save_to_file(path, 0)
    oi = oracle.index_import(cb.api._read_records(path)["index_0"])
    for ft in (1, 2):
        for q in util.planted_queries(data, 40, 62):
            ids, depth = index.search_filtered(q, 10, 0.9, ft)
            o_ids, om = oi.search(q, 10, 0.9, filter_type=ft)
            assert ids == o_ids.tolist() and depth == om["stop_depth"]
    # every variant reaches the recall target (the bound of the reference's own test, puffinn.rs:179-226: 0.8 * recall * k * queries)
    queries = util.planted_queries(data, 40, 62)
    ex = np.argsort(util.exact_distances(data, queries), axis=1)[:, :10]
    for ft in (0, 1, 2):
        hits = sum(len(set(index.search_filtered(q, 10, 0.9, ft)[0]) & set(ex[qi].tolist())) for qi, q in enumerate(queries))
        assert hits >= 0.8 * 0.9 * 10 * len(queries), (ft, hits)
    with pytest.raises(cb.api.PuffinnSearchError):
        index.search_filtered(data[0], 10, 0.9, 3)
    with pytest.raises(cb.api.PuffinnSearchError):
        index.search_filtered(data[0], 0, 0.9, 1)
    L = cl.load()
    assert L.clann_puffinn_search(None, None, 1, 0.9, 0.0, 0, None, None, None) == cl.ERR_ARG
    oi.free()
