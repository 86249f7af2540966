"""Shared helpers for the tests: synthetic data of the shapes BASELINE.json names, and comparison utilities."""
from __future__ import annotations

import numpy as np


def uniform_sphere(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def planted(n, d, seed, n_centers=None, sigma=0.6):
    """mixture of floor(0.4 sqrt(n)) Gaussian blobs, normalised to the sphere."""
    rng = np.random.default_rng(seed)
    c0 = n_centers or max(1, int(0.4 * np.sqrt(n)))
    centers = rng.standard_normal((c0, d)).astype(np.float32)
    which = rng.integers(0, c0, n)
    x = centers[which] + sigma * rng.standard_normal((n, d)).astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def planted_queries(data, nq, seed, noise=0.05):
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, data.shape[0], nq)
    q = data[pick] + noise * rng.standard_normal((nq, data.shape[1])).astype(np.float32)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


def exact_distances(data, queries):
    dn = np.linalg.norm(data, axis=1)
    qn = np.linalg.norm(queries, axis=1)
    return 1.0 - (queries @ data.T) / (qn[:, None] * dn[None, :])


def recall_at_k(data, queries, dists, counts, k):
    """src/utils/mod.rs:59-95"""
    ex = np.sort(exact_distances(data, queries), axis=1)[:, :k]
    hit = 0
    for i in range(queries.shape[0]):
        t = ex[i, k - 1] + 1e-3
        hit += int(np.sum(dists[i, : counts[i]] <= t))
    return hit / (queries.shape[0] * k)


def same_ids_up_to_ties(ids_a, dists_a, ids_b, dists_b, tol=1e-5):
    """north_star: returned top-k ids identical except for ties within 1e-5 cosine."""
    if len(ids_a) != len(ids_b):
        return False
    if list(ids_a) == list(ids_b):
        return True
    da, db = np.asarray(dists_a, np.float64), np.asarray(dists_b, np.float64)
    if not np.allclose(np.sort(da), np.sort(db), atol=tol, rtol=0):
        return False
    # ids may differ only where the distance is tied with the boundary
    only_a = set(ids_a) - set(ids_b)
    only_b = set(ids_b) - set(ids_a)
    if not only_a and not only_b:
        return True
    worst = max(da.max(), db.max())
    ok_a = all(abs(da[list(ids_a).index(i)] - worst) <= tol for i in only_a)
    ok_b = all(abs(db[list(ids_b).index(i)] - worst) <= tol for i in only_b)
    return ok_a and ok_b
