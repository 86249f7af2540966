#!/usr/bin/env python
"""bench.py — the headline benchmark of the CLANN hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...] [--dist planted|uniform]

A "step" is one search pass of the hot path over one batch of synthetic queries against a built index.
Workload at N=1 (BASELINE.json configs[2], the one the metric is quoted on): glove-100-angular shape, synthetic
1,183,514 x 100 unit vectors, 10,000 queries, num_tables=84, num_clusters_factor=0.4, k=10, delta=0.9.

Prints ONE JSON line (rank 0).
  N = 1   `value` = queries/s with queries and outputs resident in HBM, K steps issued back to back through
          clann_search_device_async (three batches in flight, four distinct query batches in rotation);
          `value_stream_ordered` = one batch at a time (clann_search_device).
  N > 1   `value` = the cluster-sharded search north_star names (clann_search_sharded: every GPU owns a share of the clusters
          and builds only those; the global batch of N x 10,000 queries is routed by nearest cluster; one all-reduce of bounds
          and one all-gather of top-k lists over NCCL) — weak scaling, the index shrinks per GPU as N grows. `replicas` = the
          other way to use N GPUs when the index fits one (index replicated, queries sharded, no collective), for comparison.
  `e2e`   the same with HOST buffers: H2D of the step's queries and D2H of its results inside the timed region.
  `roofline` = algorithmic bytes of the rerank / filter kernels over their CUDA-event duration against the measured HBM peak.
  `cpu_baseline` = the reference's own CPU implementation (oracle/_ref, real PUFFINN headers) on a bounded query sample.
`--impl reference` prints the reference arm's line (rank 0 only; other ranks exit 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec @ recall@10>=0.9 (glove-100 shape)"
UNIT = "queries/s"
# DRAM bytes (read + write) of the roofline kernels of one step from the committed ncu capture (profiles/), per workload
MEASURED_TRAFFIC = {"glove100": None}
TRAFFIC_SOURCE = "profiles/README.md"
try:
    _t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    MEASURED_TRAFFIC.update(_t.get("bytes", {}))
    TRAFFIC_SOURCE = _t.get("source", TRAFFIC_SOURCE)
except Exception:
    pass
N_QUERY_BATCHES = 4   # distinct query batches rotated through the timed loops (do not resubmit one buffer)


# ------------------------------------------------------------------------------------------------ workload

def workload(args):
    if args.small:
        return dict(name="small smoke shape (NOT the headline config)", n=100_000, d=100, nq=2_000, L=84, factor=0.4, k=10, delta=0.9)
    if args.workload == "glove25":   # BASELINE.json configs[1]
        return dict(name="glove-25-angular shape: synthetic 1,183,514x25 unit vectors, 10k queries, k=10, delta=0.9",
                    n=1_183_514, d=25, nq=10_000, L=84, factor=0.4, k=10, delta=0.9)
    if args.workload == "readme":    # BASELINE.json configs[0]
        return dict(name="README example: synthetic 10,000x128 unit vectors, 10k queries, num_tables=84, k=10, delta=0.9",
                    n=10_000, d=128, nq=10_000, L=84, factor=0.4, k=10, delta=0.9)
    if args.workload == "deep96":    # BASELINE.json configs[3]
        return dict(name="deep-image-96-angular shape: synthetic 10,000,000x96 unit vectors, 10k queries, k=10, delta=0.9",
                    n=10_000_000, d=96, nq=10_000, L=84, factor=0.4, k=10, delta=0.9)
    if args.workload == "sweep100m":  # BASELINE.json configs[4]: fp16 rows, 100k-query batches, k=100, delta from --delta
        return dict(name=f"large-scale sweep: synthetic {args.rows or 100_000_000:,}x96 fp16 unit vectors, 100k-query batches, k=100, "
                         f"delta={args.delta or 0.9}",
                    n=args.rows or 100_000_000, d=96, nq=100_000, L=84, factor=0.4, k=100, delta=args.delta or 0.9, fp16=True)
    return dict(name="glove-100-angular shape: synthetic 1,183,514x100 unit vectors, 10k queries, k=10, delta=0.9",
                n=1_183_514, d=100, nq=10_000, L=84, factor=0.4, k=10, delta=0.9)


def make_queries(data, nq, d, dist, seed):
    rq = np.random.default_rng(seed)
    n = data.shape[0]
    if dist == "planted":
        src = rq.integers(0, n, nq)
        q = data[src].astype(np.float32) + np.float32(0.05) * rq.standard_normal((nq, d), dtype=np.float32)
    else:
        src = np.full(nq, -1)
        q = rq.standard_normal((nq, d), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(q, np.float32), src


def make_data(w, dist):
    """P = planted mixture of floor(0.4 sqrt(n)) blobs (sigma 0.6), queries = point + 5% noise;
    U = i.i.d. Gaussian directions. Seeds: data 42, queries 43."""
    n, d, nq = w["n"], w["d"], w["nq"]
    rng = np.random.default_rng(42)
    if dist == "planted":
        c0 = max(1, int(0.4 * np.sqrt(n)))
        centers = rng.standard_normal((c0, d), dtype=np.float32)
        which = rng.integers(0, c0, n)
        data = centers[which]
        data += np.float32(0.6) * rng.standard_normal((n, d), dtype=np.float32)
    else:
        data = rng.standard_normal((n, d), dtype=np.float32)
    data /= np.linalg.norm(data, axis=1, keepdims=True)
    data = np.ascontiguousarray(data, np.float32)
    q, src = make_queries(data, nq, d, dist, 43)
    return data, q, src


def make_data_device(w, dist, dev, fp16):
    """The same two distributions generated on the device (torch, fixed seeds: every rank of a job draws identical rows) for the
    shapes that are too large to draw with numpy on the host in reasonable time (10M x 96, 100M x 96). Rows are normalised in
    fp32 and, for the fp16 workload, rounded to half precision — the index is then built from exactly those half values."""
    import torch
    n, d, nq = w["n"], w["d"], w["nq"]
    gen = torch.Generator(device=dev)
    gen.manual_seed(42)
    out = torch.empty((n, d), dtype=torch.float16 if fp16 else torch.float32, device=dev)
    c0 = max(1, int(0.4 * np.sqrt(n)))
    centers = torch.randn((c0, d), generator=gen, device=dev) if dist == "planted" else None
    step = 2_000_000
    for s0 in range(0, n, step):
        m = min(step, n - s0)
        x = torch.randn((m, d), generator=gen, device=dev)
        if dist == "planted":
            which = torch.randint(0, c0, (m,), generator=gen, device=dev)
            x = centers[which] + 0.6 * x
        x = x / x.norm(dim=1, keepdim=True)
        out[s0:s0 + m] = x.to(out.dtype)
    return out


def make_queries_device(data_t, nq, dist, seed, dev):
    import torch
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    n, d = data_t.shape
    if dist == "planted":
        src = torch.randint(0, n, (nq,), generator=gen, device=dev)
        q = data_t[src].float() + 0.05 * torch.randn((nq, d), generator=gen, device=dev)
    else:
        q = torch.randn((nq, d), generator=gen, device=dev)
    return (q / q.norm(dim=1, keepdim=True)).contiguous()


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md, the clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm

def reference_sample(queries, src, assignment, n_centers, n_clusters, per_cluster):
    """Bounded sample for the CPU arms: queries grouped by the cluster of their source point — `n_clusters` home clusters,
    up to `per_cluster` distinct queries each — so that the reference (2.2-2.5 s of single-threaded Monte-Carlo per PUFFINN
    index) only builds what the sample visits."""
    rng = np.random.default_rng(7)
    if src[0] >= 0:
        home = assignment[src]
        sizes = np.bincount(assignment.astype(np.int64), minlength=n_centers)
        counts = np.bincount(home.astype(np.int64), minlength=n_centers)
        eligible = [c for c in np.unique(home) if sizes[c] >= 100 and counts[c] >= min(per_cluster, 8)]
        chosen = rng.permutation(eligible)[:n_clusters]
        groups = [np.nonzero(home == c)[0][:per_cluster] for c in chosen]
        desc = (f"{sum(len(g) for g in groups)} distinct queries of the step's {len(queries)}: up to {per_cluster} per home cluster "
                f"for {len(chosen)} random clusters, one pass each")
    else:
        groups = [np.arange(min(2, len(queries)))]
        desc = f"{len(groups[0])} of {len(queries)} queries (uniform data visits every cluster)"
    return groups, desc


def _reference_worker(tmp, wi):
    """One CPU worker of the reference arm (a fresh process: no CUDA context, no inherited OpenMP pool): loads the shared arrays,
    answers its deal of queries with the real PUFFINN headers, prints one JSON object."""
    from oracle.pyoracle import RefLib
    meta = json.load(open(os.path.join(tmp, "meta.json")))
    data = np.load(os.path.join(tmp, "data.npy"), mmap_mode="r")
    data = np.ascontiguousarray(data)
    queries = np.load(os.path.join(tmp, "queries.npy"))
    centers = np.load(os.path.join(tmp, "centers.npy"))
    assignment = np.load(os.path.join(tmp, "assignment.npy"))
    radii = np.load(os.path.join(tmp, "radii.npy"))
    idx = np.load(os.path.join(tmp, f"deal_{wi}.npy"))
    eng = RefLib().clann(data, meta["L"], meta["k"], meta["delta"], centers, assignment, radii, seed_base=1234)
    for i in idx:                       # pass 1: builds what the queries visit (untimed; in the reference this is `build`)
        eng.search(queries[i])
    build_s = eng.build_seconds
    ts = time.time()
    out = []
    for i in idx:                       # pass 2: the timed pass, every query once
        ids, dd, _, ctr = eng.search(queries[i])
        out.append((int(i), dd.tolist(), ctr["visited"], ctr["distance_computations"]))
    print(json.dumps({"search_s": time.time() - ts, "build_s": build_s, "res": out}))


def run_reference(w, data, queries, src, centers, assignment, radii, workers=None, n_clusters=50, per_cluster=24):
    """The reference's own CPU implementation of the path (oracle/_ref: the real PUFFINN headers under the CLANN loop of
    index.rs:311-439) on a bounded sample. The sample's home clusters are dealt to `workers` processes (one per host core, one
    OpenMP thread each); each builds the PUFFINN indices its queries visit (lazily, in an untimed first pass), then answers its
    queries in ONE pass, one query at a time (collection.hpp:104-113 allows nothing else per index). Reported: q/s of one core
    (queries / sum of the workers' search times) and of the whole host (queries / slowest worker), recall of the sample."""
    import shutil
    import tempfile
    from oracle.pyoracle import RefLib
    if not RefLib.available():
        raise RuntimeError("oracle/_ref/libpuffinn_ref.so is missing; build it with `make -C oracle` where /root/reference exists")
    groups, desc = reference_sample(queries, src, assignment, len(centers), n_clusters, per_cluster)
    ncores = os.cpu_count() or 1
    P = max(1, min(workers or ncores, len(groups)))
    k = w["k"]
    t0 = time.time()
    tmp = tempfile.mkdtemp(prefix="clann_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        np.save(os.path.join(tmp, "data.npy"), data)
        np.save(os.path.join(tmp, "queries.npy"), queries)
        np.save(os.path.join(tmp, "centers.npy"), np.asarray(centers, np.uint64))
        np.save(os.path.join(tmp, "assignment.npy"), np.asarray(assignment, np.uint64))
        np.save(os.path.join(tmp, "radii.npy"), np.asarray(radii, np.float32))
        json.dump({"L": w["L"], "k": k, "delta": w["delta"]}, open(os.path.join(tmp, "meta.json"), "w"))
        for wi in range(P):
            deal = groups[wi::P]
            np.save(os.path.join(tmp, f"deal_{wi}.npy"), np.concatenate(deal) if deal else np.zeros(0, np.int64))
        env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-worker", tmp, str(wi)], stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, env=env, text=True) for wi in range(P)]
        results = []
        for p in procs:
            out, _ = p.communicate(timeout=900)
            lines = [l for l in out.splitlines() if l.startswith("{")]
            if lines:
                results.append(json.loads(lines[-1]))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    n_sample = sum(len(x["res"]) for x in results)
    if n_sample == 0:
        raise RuntimeError("the reference workers returned nothing")
    sum_s = sum(x["search_s"] for x in results)
    max_s = max(x["search_s"] for x in results)
    # recall of the sample against exact fp32 neighbours (utils/mod.rs:59-95)
    rows = [r for x in results for r in x["res"]]
    idx = np.array([r[0] for r in rows])
    hit = 0
    for s0 in range(0, len(idx), 64):
        sel = idx[s0:s0 + 64]
        ex = 1.0 - queries[sel] @ data.T
        kth = np.partition(ex, k - 1, axis=1)[:, k - 1]
        for j, gi in enumerate(range(s0, min(s0 + 64, len(idx)))):
            hit += int(np.sum(np.asarray(rows[gi][1], np.float32) <= kth[j] + 1e-3))
    return dict(kind="reference", sample=desc, n_sample=n_sample, build_s=max(x["build_s"] for x in results),
                qps_1thread=n_sample / sum_s, qps_allcores=n_sample / max_s, cores=P, recall=hit / (n_sample * k),
                visited=float(np.mean([r[2] for r in rows])), distcomp=float(np.mean([r[3] for r in rows])), wall_s=time.time() - t0)


# ------------------------------------------------------------------------------------------------ our arm

def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--ref-worker":
        _reference_worker(sys.argv[2], int(sys.argv[3]))
        return 0
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dist", default="planted", choices=["planted", "uniform"])
    ap.add_argument("--small", action="store_true", help="reduced shape for smoke runs; NOT a valid bench number")
    ap.add_argument("--queries", type=int, default=0, help="queries per step and GPU (default: the workload's own batch)")
    ap.add_argument("--workload", default="glove100", choices=["glove100", "glove25", "readme", "deep96", "sweep100m"],
                    help="glove100 (default) is the configuration the metric is quoted on; the others are BASELINE.json's other configs")
    ap.add_argument("--rows", type=int, default=0, help="sweep100m: number of rows (default 100 000 000)")
    ap.add_argument("--delta", type=float, default=0.0, help="sweep100m: target recall (0.8 / 0.9 / 0.95)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="N = 1: time `value` with stream-ordered calls")
    ap.add_argument("--shard", default="clusters", choices=["clusters", "replicas", "stepping"],
                    help="N > 1: clusters (default) = clann_search_sharded; replicas = index replicated, queries sharded, no "
                         "collective; stepping = the exact hand-over protocol (clann_search_begin/step/merge/end)")
    ap.add_argument("--sharded-in-flight", type=int, default=2, help="N > 1, clusters: global batches in flight per call (1..4)")
    ap.add_argument("--sharded-pipeline", choices=["stream", "multi"], default="stream",
                    help="N > 1, clusters: stream = clann_search_sharded_submit/_flush (a software pipeline across the steps, four "
                         "batches in staggered phases); multi = clann_search_sharded_multi (--sharded-in-flight batches per call, in step)")
    ap.add_argument("--no-replicas", action="store_true", help="N > 1: skip the replica comparison run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0

    w = workload(args)
    if w.get("fp16") and args.queries == 0:
        w["nq"] = max(1, w["nq"] // world)   # the sweep's batch is 100 000 queries for the whole job
    if args.queries > 0:
        w["nq"] = args.queries
        w["name"] += f" [batch overridden: {args.queries} queries per step]"
    cfg_json = {"workload": w["name"], "distribution": args.dist, "n": w["n"], "d": w["d"], "queries_per_step": w["nq"],
                "num_tables": w["L"], "num_clusters_factor": w["factor"], "k": w["k"], "delta": w["delta"],
                "l2": "index working set (2.4 GB: Q15 rows, sketches, tables) >> 126 MB L2; four distinct query batches in rotation; no explicit flush"}

    if args.impl == "reference":
        # The reference arm never touches the GPU or libclann_b200: data and clustering on the host (the clustering is the
        # oracle's C restatement of gmm.rs — the Rust crate cannot be built here — and is untimed setup), then the real
        # PUFFINN headers (oracle/_ref) run the CLANN search loop on every host core.
        from oracle.pyoracle import OracleLib
        data, queries, src = make_data(w, args.dist)
        K = max(1, int(np.floor(float(np.float32(w["factor"])) * np.sqrt(w["n"]))))
        t0 = time.time()
        centers, assignment, radii = OracleLib().gmm(data, K)
        gmm_s = time.time() - t0
        ref = run_reference(w, data, queries, src, np.asarray(centers, np.uint64), np.asarray(assignment, np.uint64),
                            np.asarray(radii, np.float32))
        line = {"impl": "reference", "metric": METRIC, "value": ref["qps_allcores"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * w["nq"] / ref["qps_allcores"],
                "ms_per_step_note": "time the host would need for one step of queries_per_step queries at the sampled rate",
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16", "data": "synthetic",
                "config": cfg_json, "recall_at_k": ref["recall"], "recall_queries_checked": ref["n_sample"],
                "cpu_baseline": {"value": ref["qps_allcores"], "unit": UNIT, "cores": ref["cores"], "kind": ref["kind"],
                                 "sample": ref["sample"], "qps_1thread": ref["qps_1thread"], "recall_at_k": ref["recall"],
                                 "clusters_visited_per_query": ref["visited"], "distance_computations_per_query": ref["distcomp"],
                                 "index_build_s_for_sample": ref["build_s"], "clustering_s": gmm_s,
                                 "clustering": "oracle C restatement of gmm.rs on the host (no Rust toolchain), untimed setup"},
                "e2e": {"value": ref["qps_allcores"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import clann_b200 as cb
    from clann_b200 import _lib as cl
    from clann_b200.distributed import ClusterShardedSearcher, ShardedSearcher

    dev = torch.device("cuda", local_rank)
    nq, k, d = w["nq"], w["k"], w["d"]
    big = w["n"] >= 5_000_000 or w.get("fp16", False)   # rows generated (and kept) on the device
    data_t = None
    if big:
        data_t = make_data_device(w, args.dist, dev, w.get("fp16", False))
        data, queries, src = None, make_queries_device(data_t, nq, args.dist, 43, dev).cpu().numpy(), np.full(nq, -1)
    else:
        data, queries, src = make_data(w, args.dist)
    mode = "single" if world == 1 else args.shard
    gnq = nq * world if mode in ("clusters", "stepping") else nq      # queries every rank holds per step
    global_nq = nq * world if world > 1 else nq                        # queries the whole job answers per step
    # distinct query batches (host): batch j of the job; in replica mode every rank draws its own
    seed0 = 43 + (1000 * rank if mode == "replicas" else 0)
    batches = []
    for j in range(N_QUERY_BATCHES):
        if j == 0 and gnq == nq and mode != "replicas":
            batches.append(queries)
        elif big:
            batches.append(make_queries_device(data_t, gnq, args.dist, seed0 + 17 * j, dev).cpu().numpy())
        else:
            batches.append(make_queries(data, gnq, d, args.dist, seed0 + 17 * j)[0])
    parallelism = ("single GPU" if world == 1 else
                   f"clusters sharded over {world} GPUs (each builds and holds its own clusters only); global batch {global_nq} queries per step "
                   f"routed by nearest cluster; all-gather of nearest-cluster ids, all-reduce(min) of bounds, all-gather of top-k lists (NCCL)"
                   if mode == "clusters" else
                   f"clusters sharded over {world} GPUs, exact stepping: all-gather of the per-query states per step" if mode == "stepping" else
                   f"index replicated on {world} GPUs, {nq} queries per GPU per step (global batch {global_nq}), no data-path collective")

    # ---- build (untimed setup of the search benchmark; reported on its own)
    def build_index(sharded):
        t0 = time.time()
        conf = cb.Config(w["L"], w["factor"], w["k"], w["delta"], "bench")
        if big:
            ix = cb.ClusteredIndex.from_rows(conf, data_t.data_ptr(), w["n"], d, "f16" if w.get("fp16") else "f32", on_device=True)
        else:
            ix = cb.init_with_config(data, conf)
        ix.set_option("seed", 1234)
        srch = None
        if sharded:
            ix.set_option("shard_count", world)
            ix.set_option("shard_rank", rank)
            if mode == "clusters":
                # the communicator exists before the build: greedy k-center then runs sharded too (each rank its rows, one 8-byte
                # all-reduce of the arg-max key per pass)
                srch = ClusterShardedSearcher(ix, world, rank)
                barrier0()
                t0 = time.time()
        ix.build()
        torch.cuda.synchronize()
        return ix, time.time() - t0, ix.export(cl.X_BUILD_MS, 0, np.float64).copy(), srch

    def barrier0():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    index, build_wall, build_ms, pre_searcher = build_index(mode in ("clusters", "stepping"))
    if world == 1 and not big:   # a second, warm build: the first one pays module loading and the first cudaMallocs
        index.close()
        index, build_wall2, build_ms2, _ = build_index(False)
    else:
        build_wall2, build_ms2 = build_wall, build_ms
    K = index.num_clusters
    if not big:   # only the CPU baseline needs the clustering on the host
        centers = index.export(cl.X_CENTERS, 0, np.uint64).copy()
        assignment = index.export(cl.X_ASSIGNMENT, 0, np.uint64).copy()
        radii = index.export(cl.X_RADII, 0, np.float32).copy()

    d_batches = [torch.from_numpy(b).to(dev) for b in batches]
    d_q = d_batches[0]
    d_ids = torch.empty((gnq, k), dtype=torch.int32, device=dev)
    d_dists = torch.empty((gnq, k), dtype=torch.float32, device=dev)
    d_counts = torch.empty(gnq, dtype=torch.int32, device=dev)
    if mode == "clusters":
        searcher = pre_searcher
    elif mode == "stepping":
        searcher = ShardedSearcher(index, world, rank)
    else:
        searcher = ShardedSearcher(index, 1, 0)
    single = mode in ("single", "replicas")  # this rank runs the whole single-GPU path on its own batch

    def step_device(i=0):
        searcher.search_device(d_batches[i % N_QUERY_BATCHES], d_ids, d_dists, d_counts)

    pipelined = single and not args.no_pipeline
    # 12 output sets: calls i and i + 12 land on the same internal stream for every pipeline depth (2, 3 or 4)
    NSETS = 12
    outs = [(d_ids, d_dists, d_counts)] + [(torch.empty_like(d_ids), torch.empty_like(d_dists), torch.empty_like(d_counts)) for _ in range(NSETS - 1)]
    cur_stream = torch.cuda.current_stream().cuda_stream

    def run_steps(steps):
        if mode == "clusters" and not args.no_pipeline and args.sharded_pipeline == "stream":
            # one batch per step into the staggered pipeline: the sharded counterpart of clann_search_device_async / _flush
            for i in range(steps):
                searcher.submit(d_batches[i % N_QUERY_BATCHES], outs[i % NSETS])
            searcher.flush()
            return
        if mode == "clusters" and not args.no_pipeline:
            # several whole batches in flight (clann_search_sharded_multi), all in the same phase
            i = 0
            while i < steps:
                nb = min(args.sharded_in_flight, steps - i)
                searcher.search_device_multi([d_batches[(i + j) % N_QUERY_BATCHES] for j in range(nb)], [outs[j] for j in range(nb)])
                i += nb
            return
        if not pipelined:
            for i in range(steps):
                step_device(i)
            return
        for i in range(steps):
            o = outs[i % NSETS]
            q = d_batches[i % N_QUERY_BATCHES]
            if index._lib.clann_search_device_async(index.handle, q.data_ptr(), gnq, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) != 0:
                raise RuntimeError(cl.last_error())
        if index._lib.clann_search_flush(index.handle, cur_stream) != 0:
            raise RuntimeError(cl.last_error())

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    for i in range(args.warmup):
        step_device(i)
    # the streaming pipeline rotates four internal lanes (stream + workspaces each): every one of them is used once before the clock
    # starts, whatever W is, so that no first-use allocation lands in the timed region
    run_steps(max(args.warmup, 4) if mode == "clusters" else args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # the timed region: exactly K steps; the stream is idle at e0 (barrier + synchronize), so e0..e1 covers every kernel of
    # the K steps whichever internal stream ran it (clann_search_flush makes the current stream wait for all of them)
    total_ms = timed(run_steps, args.steps)
    ordered_ms = timed(lambda steps: [step_device(i) for i in range(steps)], args.steps)
    clocks = sampler.stop()
    # kernels per step: counted by the library on the single-GPU path; the sharded search launches 3 (route) + 4 (select, fills) +
    # 14 (round one: gather, 8 prep, 2 precompute, probe, pack, collect) + 1 + 11 (round two) + 1 (merge) of its own kernels
    launches_per_step = 34 if mode == "clusters" else getattr(searcher, "last_launches", 0)
    probe_ms = prep_ms = None
    if single:  # per-kernel split from a separate pass (events inside the library); same stream, same inputs
        ps, pr = 0.0, 0.0
        for i in range(args.steps):
            step_device(i)
            prof = index.search_profile()
            ps += prof["probe_ms"]; pr += prof["prep_ms"]
        probe_ms, prep_ms = ps / args.steps, pr / args.steps

    ms_per_step = total_ms / args.steps
    value = global_nq / (ms_per_step / 1000.0)

    # ---- end to end: host buffers in and out, copies inside the timed region
    h_batches = [torch.from_numpy(b).pin_memory() for b in batches]
    h_ids = torch.empty((gnq, k), dtype=torch.int32).pin_memory()
    h_dists = torch.empty((gnq, k), dtype=torch.float32).pin_memory()
    h_counts = torch.empty(gnq, dtype=torch.int32).pin_memory()
    lo_q, hi_q = (rank * nq, (rank + 1) * nq) if mode in ("clusters", "stepping") else (0, nq)

    def e2e_sync(steps):
        for i in range(steps):
            hq = h_batches[i % N_QUERY_BATCHES]
            if single:
                # the reference-facing call: host pointers in, host pointers out (copies inside clann_search)
                if index._lib.clann_search(index.handle, hq.data_ptr(), nq, h_ids.data_ptr(), h_dists.data_ptr(), h_counts.data_ptr()) != 0:
                    raise RuntimeError(cl.last_error())
            else:
                # every rank uploads ITS slice of the step's queries and the slices are all-gathered over NVLink (queries are
                # broadcast); after the search every rank downloads the results of its slice
                import torch.distributed as dist
                d_slice.copy_(hq[lo_q:hi_q], non_blocking=True)
                dist.all_gather_into_tensor(d_q, d_slice)
                searcher.search_device(d_q, d_ids, d_dists, d_counts)
                h_ids[lo_q:hi_q].copy_(d_ids[lo_q:hi_q], non_blocking=True)
                h_dists[lo_q:hi_q].copy_(d_dists[lo_q:hi_q], non_blocking=True)
                h_counts[lo_q:hi_q].copy_(d_counts[lo_q:hi_q], non_blocking=True)
                torch.cuda.synchronize()

    if not single:
        d_q = torch.empty_like(d_batches[0])
        d_slice = torch.empty((nq, d), dtype=torch.float32, device=dev)
    e2e_sync(2)
    e2e_sync_ms = timed(e2e_sync, args.steps) / args.steps
    e2e_ms = e2e_sync_ms
    e2e_mode = ("clann_search: one synchronous call per step (host buffers in, host buffers out)" if single else
                "per step: H2D of this rank's slice of the queries, NCCL all-gather of the slices, clann_search_sharded, D2H of the "
                "slice's results; blocking")
    if pipelined:
        # the same through clann_search_async: every step's H2D copy, search and D2H copies on its batch stream, three batches in
        # flight, each with its own pinned output buffers; clann_search_wait before the clock stops
        h_outs = [(h_ids, h_dists, h_counts)] + [(torch.empty_like(h_ids).pin_memory(), torch.empty_like(h_dists).pin_memory(),
                                                  torch.empty_like(h_counts).pin_memory()) for _ in range(NSETS - 1)]

        def run_e2e_async(steps, batch_of=lambda i: i % N_QUERY_BATCHES):
            for i in range(steps):
                o = h_outs[i % NSETS]
                hq = h_batches[batch_of(i)]
                if index._lib.clann_search_async(index.handle, hq.data_ptr(), nq, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) != 0:
                    raise RuntimeError(cl.last_error())
            if index._lib.clann_search_wait(index.handle) != 0:
                raise RuntimeError(cl.last_error())

        run_e2e_async(3)
        if index._lib.clann_search(index.handle, h_batches[0].data_ptr(), nq, h_ids.data_ptr(), h_dists.data_ptr(), h_counts.data_ptr()) != 0:
            raise RuntimeError(cl.last_error())
        want = (h_ids.clone(), h_dists.clone(), h_counts.clone())
        run_e2e_async(NSETS, batch_of=lambda i: 0)
        for o in h_outs:
            if not (torch.equal(o[0], want[0]) and torch.equal(o[1], want[1]) and torch.equal(o[2], want[2])):
                raise RuntimeError("clann_search_async returned results that differ from clann_search")
        e2e_ms = timed(run_e2e_async, args.steps) / args.steps
        e2e_mode = ("clann_search_async + clann_search_wait: host buffers in and out, H2D / search / D2H of each step on its batch "
                    "stream, three batches in flight; results identical to clann_search (checked)")
    if mode == "clusters" and not args.no_pipeline and args.sharded_pipeline == "stream":
        # the same through the streaming form: per step H2D of this rank's slice, all-gather of the slices, submit; the batch that
        # the call finished (three steps older) is downloaded behind it; flush + the last downloads before the clock stops
        SD = 6
        dq_sets = [torch.empty_like(d_batches[0]) for _ in range(SD)]
        ds_sets = [torch.empty((nq, d), dtype=torch.float32, device=dev) for _ in range(SD)]
        h_outs = [(torch.empty_like(h_ids).pin_memory(), torch.empty_like(h_dists).pin_memory(), torch.empty_like(h_counts).pin_memory())
                  for _ in range(SD)]

        def download(j):
            o, ho = outs[j % SD], h_outs[j % SD]
            for a, b in zip(ho, o):
                a[lo_q:hi_q].copy_(b[lo_q:hi_q], non_blocking=True)

        def run_e2e_stream(steps, batch_of=lambda i: i % N_QUERY_BATCHES):
            import torch.distributed as dist
            for i in range(steps):
                ds_sets[i % SD].copy_(h_batches[batch_of(i)][lo_q:hi_q], non_blocking=True)
                dist.all_gather_into_tensor(dq_sets[i % SD], ds_sets[i % SD])
                searcher.submit(dq_sets[i % SD], outs[i % SD])
                if i >= 3:
                    download(i - 3)
            searcher.flush()
            for j in range(max(0, steps - 3), steps):
                download(j)
            torch.cuda.synchronize()

        run_e2e_stream(SD, batch_of=lambda i: 0)
        searcher.search_device(d_batches[0], d_ids, d_dists, d_counts)
        torch.cuda.synchronize()
        for ho in h_outs:
            if not (torch.equal(ho[0][lo_q:hi_q], d_ids[lo_q:hi_q].cpu()) and torch.equal(ho[1][lo_q:hi_q], d_dists[lo_q:hi_q].cpu())):
                raise RuntimeError("clann_search_sharded_submit returned results that differ from clann_search_sharded")
        e2e_ms = timed(run_e2e_stream, args.steps) / args.steps
        e2e_mode = ("per step: H2D of this rank's slice of the queries, NCCL all-gather of the slices, clann_search_sharded_submit; "
                    "D2H of the slice's results of the batch the call finished; clann_search_sharded_flush and the last downloads "
                    "before the clock stops; results identical to clann_search_sharded (checked)")
    e2e_value = global_nq / (e2e_ms / 1000.0)

    # ---- correctness of what was timed: the pipelined batches return what the stream-ordered call returns
    if pipelined:
        for i in range(3):
            o = outs[i]
            if index._lib.clann_search_device_async(index.handle, d_batches[1].data_ptr(), gnq, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) != 0:
                raise RuntimeError(cl.last_error())
        index._lib.clann_search_flush(index.handle, cur_stream)
        torch.cuda.synchronize()
        a_ids, a_dists = outs[1][0].clone(), outs[1][1].clone()
        b_ids, b_dists = outs[2][0].clone(), outs[2][1].clone()
        step_device(1)
        torch.cuda.synchronize()
        if not (torch.equal(a_ids, d_ids) and torch.equal(b_ids, d_ids) and torch.equal(a_dists, d_dists) and torch.equal(b_dists, d_dists)):
            raise RuntimeError("pipelined batches returned results that differ from the stream-ordered call")

    # ---- recall@k against exact fp32 neighbours (utils/mod.rs:59-95), on this rank's share of batch 0
    step_device(0)
    torch.cuda.synchronize()
    ids = d_ids.cpu().numpy().view(np.uint32); dists = d_dists.cpu().numpy(); counts = d_counts.cpu().numpy()
    nchk = min(gnq, 2000)
    with torch.no_grad():
        dd = data_t if big else torch.from_numpy(data).to(dev)
        dq0 = d_batches[0][:nchk]
        best = torch.full((nchk, k), -2.0, device=dev)
        rows_step = 2_000_000
        qs = max(8, min(nchk, int(2.5e8 // min(w["n"], rows_step))))   # keep the similarity tile under ~1 GB
        for r0 in range(0, w["n"], rows_step):   # running top-k over row chunks (the big shapes do not fit one tile)
            blk = dd[r0:r0 + rows_step].float()
            for s in range(0, nchk, qs):
                e = min(s + qs, nchk)
                sim = dq0[s:e] @ blk.T
                cand = torch.cat([best[s:e], torch.topk(sim, min(k, sim.shape[1]), dim=1).values], dim=1)
                best[s:e] = torch.topk(cand, k, dim=1).values
        kth = (1.0 - best[:, k - 1]).cpu().numpy()
        if not big:
            del dd
    hit = sum(int(np.sum(dists[i, :counts[i]] <= kth[i] + 1e-3)) for i in range(nchk))
    recall = hit / (nchk * k)
    if args.dist == "planted" and not args.small and recall < 0.9:
        raise RuntimeError(f"recall@{k} = {recall:.4f} < 0.9: the metric is defined at recall >= 0.9, refusing to report a throughput")

    # ---- per-query work (this rank's share in sharded mode, summed over the ranks below)
    if mode == "clusters":
        ctr = index.counters(gnq)
        tot = torch.tensor([float(ctr["candidates"].sum()), float(ctr["distance_computations"].sum()), float(ctr["clusters_visited"].sum())],
                           dtype=torch.float64, device=dev)
        import torch.distributed as dist
        dist.all_reduce(tot)
        cand, dc, vis = (float(x) for x in tot.tolist())
        routed, still_open = searcher.stats()
    else:
        ctr = index.counters(gnq) if single else searcher.counters(gnq)
        cand = float(ctr["candidates"].sum()); dc = float(ctr["distance_computations"].sum()); vis = float(ctr["clusters_visited"].sum())
        routed = still_open = None

    # ---- N > 1: the replica arrangement for comparison (index replicated, every rank its own batch, no collective)
    replicas = None
    if world > 1 and mode == "clusters" and not args.no_replicas and w["n"] <= 20_000_000:
        rindex, _, _, _ = build_index(False)
        rq = (make_queries_device(data_t, nq, args.dist, 43 + 1000 * rank, dev) if big else
              torch.from_numpy(make_queries(data, nq, d, args.dist, 43 + 1000 * rank)[0]).to(dev))
        r_outs = [(torch.empty((nq, k), dtype=torch.int32, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev),
                   torch.empty(nq, dtype=torch.int32, device=dev)) for _ in range(NSETS)]

        def run_replica(steps):
            for i in range(steps):
                o = r_outs[i % NSETS]
                if rindex._lib.clann_search_device_async(rindex.handle, rq.data_ptr(), nq, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) != 0:
                    raise RuntimeError(cl.last_error())
            rindex._lib.clann_search_flush(rindex.handle, cur_stream)

        run_replica(args.warmup)
        r_ms = timed(run_replica, args.steps) / args.steps
        replicas = {"value": nq * world / (r_ms / 1000.0), "unit": UNIT, "ms_per_step": r_ms,
                    "what": f"index replicated on {world} GPUs, {nq} queries per GPU per step, three batches in flight, no collective"}
        rindex.close()

    sweep = None
    if w.get("fp16") and not args.delta:
        sweep = []
        for dl in (0.8, 0.9, 0.95):
            index.set_delta(dl)
            for i in range(2):
                step_device(i)
            ms = timed(lambda steps: [step_device(i) for i in range(steps)], max(3, args.steps // 2)) / max(3, args.steps // 2)
            step_device(0)
            torch.cuda.synchronize()
            dd_ = d_dists.cpu().numpy(); cc_ = d_counts.cpu().numpy()
            hit_ = sum(int(np.sum(dd_[i, :cc_[i]] <= kth[i] + 1e-3)) for i in range(nchk))
            sweep.append({"delta": dl, "value": global_nq / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "recall_at_k": hit_ / (nchk * k)})
        index.set_delta(w["delta"])

    if rank == 0:
        sl = (d + 15) // 16 * 16
        rerank_bytes = dc * 2 * sl + vis * k * 4 * d
        filter_bytes = cand * 12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = None
        if probe_ms:
            ach = (rerank_bytes + filter_bytes) / (probe_ms / 1000.0) / 1e9
            roofline = {"bound": "hbm", "kernel": "k_dense_sims + k_first_ranges + k_probe (rerank and anchors of first visits streamed, then one warp per query)",
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                        "traffic": MEASURED_TRAFFIC.get(args.workload if not args.small else "small"), "traffic_source": TRAFFIC_SOURCE,
                        "kernel_ms": probe_ms, "prep_ms": prep_ms,
                        "kernel_ms_note": "k_dense_sims (the rerank arithmetic of every query's first visit, streamed) + k_first_ranges (its "
                                          "anchors) + k_probe: the kernels between the library's events, all counted against the same "
                                          "algorithmic bytes",
                        "algorithmic_bytes_per_launch": rerank_bytes + filter_bytes,
                        "rerank_gbs": rerank_bytes / (probe_ms / 1000.0) / 1e9, "filter_gbs": filter_bytes / (probe_ms / 1000.0) / 1e9}
        # build roofline: the greedy k-center passes stream n x d fp32 once per centre
        gmm_bytes = float(K) * w["n"] * d * 4
        pass_ms = float(build_ms2[4]) if len(build_ms2) > 4 and build_ms2[4] > 0 else float(build_ms2[0])
        build_roofline = {"bound": "hbm", "kernel": "k_gmm_pass_v x K (one pass over the rows per centre; rows the triangle inequality rules out are not read)",
                          "achieved": gmm_bytes / (pass_ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": gmm_bytes / (pass_ms / 1000.0) / 1e9 / peak, "kernel_ms": pass_ms,
                          "note": "algorithmic bytes K x n x d x 4 over the K passes (CUDA events around the pass loop: k_gmm_centre_dists + "
                                  "k_gmm_pass_v per centre); gmm_ms below is the whole clustering phase incl. assignment inversion and host hand-offs"}
        cpu = None
        if not args.no_cpu_baseline and world == 1 and args.workload in ("glove100", "glove25", "readme"):
            try:
                ref = run_reference(w, data, queries, src, centers, assignment, radii, n_clusters=12, per_cluster=24)
                cpu = {"value": ref["qps_1thread"], "unit": UNIT, "cores": 1, "kind": ref["kind"], "sample": ref["sample"],
                       "qps_allcores": ref["qps_allcores"], "cores_allcores": ref["cores"], "recall_at_k": ref["recall"],
                       "clusters_visited_per_query": ref["visited"], "distance_computations_per_query": ref["distcomp"],
                       "index_build_s_for_sample": ref["build_s"]}
            except Exception as e:  # the baseline is reported, never required for our own number
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i16",
            "value_stream_ordered": global_nq / (ordered_ms / args.steps / 1000.0), "ms_per_step_stream_ordered": ordered_ms / args.steps,
            "pipeline": ("clann_search_device_async: three batches in flight on three internal streams; outputs identical to the "
                         "stream-ordered call (checked)") if pipelined else
                        (("clann_search_sharded_submit / _flush: one global batch per step into a software pipeline across the steps "
                          "(route | round one | open-query scoring | round two + merge of four consecutive batches overlap)"
                          if args.sharded_pipeline == "stream" else
                          f"clann_search_sharded_multi: {args.sharded_in_flight} global batches in flight, their phases interleaved on internal streams")
                         if mode == "clusters" and not args.no_pipeline else "none (stream-ordered calls)"),
            "data": "synthetic", "config": cfg_json, "parallelism": parallelism, "recall_at_k": recall, "recall_queries_checked": nchk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": global_nq * d * 4,
                    "d2h_bytes_per_step": global_nq * k * 8 + global_nq * 4,
                    "ms_per_step": e2e_ms, "call": e2e_mode,
                    "value_synchronous_call": global_nq / (e2e_sync_ms / 1000.0), "ms_per_step_synchronous_call": e2e_sync_ms},
            "gpu_launches": int(max(launches_per_step, 9) * args.steps * (world if mode == "replicas" else 1)),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "build": {"wall_s": build_wall2, "gmm_ms": build_ms2[0], "hash_ms": build_ms2[1], "sort_ms": build_ms2[2], "device_ms": build_ms2[3],
                      "clusters": int(K), "first_build_wall_s": build_wall, "first_build_device_ms": build_ms[3], "roofline": build_roofline,
                      "note": "second (warm) build of the same index; the first pays module loading"
                              if world == 1 else "collective build: greedy k-center over each rank's share of the rows with an all-reduce of the "
                                                 "arg-max per pass, then the tables of this rank's clusters only"},
            "per_query": {"clusters_visited": vis / gnq, "candidates": cand / gnq, "distance_computations": dc / gnq},
        }
        if mode == "clusters":
            line["sharded"] = {"routed_to_rank0_round_one": routed, "open_after_round_one": still_open, "global_batch": gnq,
                               "phase_ms_rank0": getattr(searcher, "phase_ms", None)}
            line["replicas"] = replicas
        if sweep:
            line["delta_sweep"] = sweep
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
