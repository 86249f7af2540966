"""Generates the golden fixtures of tests/golden/ from the REAL reference (oracle/_ref, built from /root/reference).

Run here (where /root/reference exists):  python tests/golden/make_golden.py
The reference cannot travel to the GPU box, these vectors can. Each fixture holds the inputs (data, queries, RNG seed),
the reference's serialized index (functions + Q15 rows + sketches + tables) and its outputs for the same queries:
table codes, sketches, anchors, per-depth ranges, result ids and the per-query counters.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.pyoracle import RefLib, build  # noqa: E402

CASES = {
    "puffinn_d25": dict(n=400, d=25, L=8, seed=2025, nq=40),
    "puffinn_d100": dict(n=300, d=100, L=6, seed=2100, nq=30),
}
SEARCHES = [(10, 0.9, float("-inf")), (1, 0.5, 0.6), (5, 0.95, 0.75)]
# FilterType::None (1) / FilterType::Simple (2) of Index::search (collection.hpp:22-34,671-765): (k, recall)
FILTER_SEARCHES = [(10, 0.9), (1, 0.5), (5, 0.95), (3, 0.2)]


def main():
    build()
    R = RefLib()
    for name, c in CASES.items():
        rng = np.random.default_rng(c["seed"])
        centers = rng.standard_normal((4, c["d"])).astype(np.float32)
        data = centers[rng.integers(0, 4, c["n"])] + 0.7 * rng.standard_normal((c["n"], c["d"])).astype(np.float32)
        data = np.ascontiguousarray(data, np.float32)  # deliberately NOT normalised: the index normalises (unit_vector.hpp:61-89)
        queries = data[rng.integers(0, c["n"], c["nq"])] + 0.2 * rng.standard_normal((c["nq"], c["d"])).astype(np.float32)
        queries = np.ascontiguousarray(queries, np.float32)
        ix = R.index(c["d"], data, c["L"], seed=c["seed"])
        stream = np.frombuffer(ix.serialize(), np.uint8)
        codes = np.stack([ix.query_codes(q, c["L"]).astype(np.uint32) for q in queries])
        sketches = np.stack([ix.query_sketches(q) for q in queries])
        anchors, ranges = zip(*[ix.query_ranges(q, c["L"]) for q in queries])
        q15 = np.stack([ix.store_q15(q) for q in queries])
        res_ids = np.full((len(SEARCHES), c["nq"], 10), 0xFFFFFFFF, np.uint32)
        res_cnt = np.zeros((len(SEARCHES), c["nq"]), np.uint32)
        res_met = np.zeros((len(SEARCHES), c["nq"], 4), np.uint32)
        for si, (k, rec, ms) in enumerate(SEARCHES):
            for qi, q in enumerate(queries):
                ids, m = ix.search(q, k, rec, ms)
                res_ids[si, qi, : len(ids)] = ids
                res_cnt[si, qi] = len(ids)
                res_met[si, qi] = [m["distance_computations"], m["candidates"], m["hash_length"], m["considered_maps"]]
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, data=data, queries=queries, L=c["L"], seed=c["seed"], stream=stream, query_q15=q15,
                            query_codes=codes, query_sketches=sketches, anchors=np.stack(anchors), ranges=np.stack(ranges),
                            searches=np.array(SEARCHES, np.float64), res_ids=res_ids, res_cnt=res_cnt, res_met=res_met)
        print(name, os.path.getsize(out) // 1024, "KiB")
        ix.free()


def filters():
    """<name>_filters.npz: the reference's answers with FilterType::None / Simple on the index stored in <name>.npz (the stream is
    loaded, not rebuilt, so the two fixtures describe the same index): ids best first, their count, and hash_length — the depth
    at which the per-depth stop rule fired (0 = never)."""
    R = RefLib()
    for name in CASES:
        g = np.load(os.path.join(HERE, name + ".npz"))
        ix = R.index_from_stream(g["stream"].tobytes())
        nq = len(g["queries"])
        res_ids = np.full((2, len(FILTER_SEARCHES), nq, 10), 0xFFFFFFFF, np.uint32)
        res_cnt = np.zeros((2, len(FILTER_SEARCHES), nq), np.uint32)
        res_depth = np.zeros((2, len(FILTER_SEARCHES), nq), np.uint32)
        for fi, ft in enumerate((1, 2)):
            for si, (k, rec) in enumerate(FILTER_SEARCHES):
                for qi, q in enumerate(g["queries"]):
                    ids, m = ix.search(q, k, rec, filter_type=ft)
                    res_ids[fi, si, qi, : len(ids)] = ids
                    res_cnt[fi, si, qi] = len(ids)
                    res_depth[fi, si, qi] = m["hash_length"]
                    assert m["distance_computations"] == 0 and m["candidates"] == 0  # neither variant counts (only :865,904,921)
        out = os.path.join(HERE, name + "_filters.npz")
        np.savez_compressed(out, filter_types=np.array([1, 2]), searches=np.array(FILTER_SEARCHES, np.float64), res_ids=res_ids,
                            res_cnt=res_cnt, res_depth=res_depth)
        print(name + "_filters", os.path.getsize(out) // 1024, "KiB")
        ix.free()


if __name__ == "__main__":
    if sys.argv[1:] == ["filters"]:
        build()
        filters()
    else:
        main()
        filters()
