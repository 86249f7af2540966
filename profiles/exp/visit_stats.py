#!/usr/bin/env python
"""Distribution of clusters visited per query vs the work-order predictor (clusters whose ball reaches closer than the
nearest centre), glove-100 shape, planted queries. Prints a small table; used to choose the longest-first threshold."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import clann_b200 as cb
from clann_b200 import _lib as cl


class A:
    small, workload = False, "glove100"


w = bench.workload(A)
data, queries, _ = bench.make_data(w, "planted")
index = cb.init_with_config(data, cb.Config(w["L"], w["factor"], w["k"], w["delta"], "stats"))
index.set_option("seed", 1234)
index.build()
ids, dists, counts = index.search_batch(queries)
ctr = index.counters(len(queries))
vis = ctr["clusters_visited"].astype(np.int64)
cand = ctr["candidates"].astype(np.int64)
dc = ctr["distance_computations"].astype(np.int64)
centers = index.export(cl.X_CENTERS, 0, np.uint64).astype(np.int64)
radii = index.export(cl.X_RADII, 0, np.float32)
C = data[centers]
cd = 1.0 - (queries @ C.T) / (np.linalg.norm(queries, axis=1, keepdims=True) * np.linalg.norm(C, axis=1)[None, :])
nearest = cd.min(axis=1)
est = ((cd - radii[None, :]) <= nearest[:, None]).sum(axis=1)
print("visited: mean %.3f max %d; histogram" % (vis.mean(), vis.max()), np.bincount(vis)[:40])
print("est: mean %.2f max %d; histogram" % (est.mean(), est.max()), np.bincount(est)[:40])
print("candidates per query: mean %.0f p50 %.0f p90 %.0f p99 %.0f max %d" % (cand.mean(), *np.percentile(cand, [50, 90, 99]), cand.max()))
print("distcomp per query: mean %.0f p50 %.0f p90 %.0f p99 %.0f max %d" % (dc.mean(), *np.percentile(dc, [50, 90, 99]), dc.max()))
for lo, hi in ((0, 1), (2, 2), (3, 4), (5, 8), (9, 16), (17, 10**6)):
    m = (est >= lo) & (est <= hi)
    if m.any():
        print(f"est in [{lo},{hi}]: {m.sum()} queries, visited mean {vis[m].mean():.2f} max {vis[m].max()}, cand mean {cand[m].mean():.0f} max {cand[m].max()}")
print("corr(est, cand) = %.3f, corr(vis, cand) = %.3f" % (np.corrcoef(est, cand)[0, 1], np.corrcoef(vis, cand)[0, 1]))
order = np.argsort(-cand)
print("top-20 candidates:", cand[order[:20]].tolist(), "their visited:", vis[order[:20]].tolist(), "est:", est[order[:20]].tolist())
