#!/usr/bin/env python
"""bench.py — the headline benchmark of the CLANN hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dist planted|uniform] [--small]

A "step" is one search pass of the hot path over one batch of synthetic queries against a built index.
Workload at N=1 (BASELINE.json configs[2], the one the metric is quoted on): glove-100-angular shape, synthetic
1,183,514 x 100 unit vectors, 10,000 queries, num_tables=84, num_clusters_factor=0.4, k=10, delta=0.9.

Prints ONE JSON line (rank 0). `value` = queries/s with queries and outputs resident in HBM, K steps issued back to back
through clann_search_device_async (three batches in flight; `value_stream_ordered` = one batch at a time, clann_search_device),
`e2e` = the same with HOST buffers, H2D + D2H inside the timed region, through clann_search_async / clann_search_wait
(`value_synchronous_call` = one blocking clann_search per step),
`roofline` = algorithmic bytes of the probe kernel / its CUDA-event duration against the measured HBM peak,
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
# DRAM bytes of one k_probe launch from the committed ncu capture (profiles/), per workload; None = not captured
MEASURED_TRAFFIC = {"glove100": 3.57e9}


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
    if args.workload == "deep96":    # BASELINE.json configs[3] on one GPU (the index is 16 GB; sharding is optional)
        return dict(name="deep-image-96-angular shape: synthetic 10,000,000x96 unit vectors, 10k queries, k=10, delta=0.9",
                    n=10_000_000, d=96, nq=10_000, L=84, factor=0.4, k=10, delta=0.9)
    return dict(name="glove-100-angular shape: synthetic 1,183,514x100 unit vectors, 10k queries, k=10, delta=0.9",
                n=1_183_514, d=100, nq=10_000, L=84, factor=0.4, k=10, delta=0.9)


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
    rq = np.random.default_rng(43)
    if dist == "planted":
        src = rq.integers(0, n, nq)
        q = data[src] + np.float32(0.05) * rq.standard_normal((nq, d), dtype=np.float32)
    else:
        src = np.full(nq, -1)
        q = rq.standard_normal((nq, d), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(data, np.float32), np.ascontiguousarray(q, np.float32), src


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

def reference_sample(w, data, queries, src, centers, assignment, radii, n_clusters=6, per_cluster=12):
    """Bounded sample for the CPU arms: queries grouped by the cluster of their source point, a few clusters, a dozen
    queries each, so that the reference (2.2-2.5 s of Monte-Carlo per PUFFINN index) only builds what the sample visits."""
    rng = np.random.default_rng(7)
    if src[0] >= 0:
        home = assignment[src]
        sizes = np.bincount(assignment.astype(np.int64), minlength=len(centers))
        eligible = [c for c in np.unique(home) if sizes[c] >= 100]
        chosen = rng.permutation(eligible)[:n_clusters]
        idx = np.concatenate([np.nonzero(home == c)[0][:per_cluster] for c in chosen])
        desc = f"{len(idx)} of {len(queries)} queries: {per_cluster} per home cluster for {len(chosen)} random clusters"
    else:
        idx = np.arange(min(2, len(queries)))
        desc = f"{len(idx)} of {len(queries)} queries (uniform data visits every cluster)"
    return idx, desc


def run_reference(w, data, queries, src, centers, assignment, radii, threads=None):
    """The reference's own CPU implementation of the path (oracle/_ref: the real PUFFINN headers + the CLANN loop) on a
    bounded sample. Index builds are lazy and excluded from the search time. Search is one query at a time per worker, as
    in the reference (collection.hpp:104-113); `threads` forked workers each own a copy-on-write view of the indices."""
    from oracle.pyoracle import RefLib, OracleLib
    idx, desc = reference_sample(w, data, queries, src, centers, assignment, radii)
    sample = queries[idx]
    kind = "reference" if RefLib.available() else "port"
    if kind == "reference":
        eng = RefLib().clann(data, w["L"], w["k"], w["delta"], centers, assignment, radii, seed_base=1234)
    else:
        raise RuntimeError("oracle/_ref/libpuffinn_ref.so is missing; build it with `make -C oracle` where /root/reference exists")
    t0 = time.time()
    results = [eng.search(q) for q in sample]          # first pass builds the visited clusters (lazy), untimed
    build_s = eng.build_seconds
    ncores = os.cpu_count() or 1
    P = threads or max(1, min(ncores, len(sample)))
    # timed pass: P forked workers, each its slice, sequential queries inside a worker
    eng.search_seconds.value = 0.0
    t1 = time.time()
    for q in sample:
        eng.search(q)
    single_s = time.time() - t1
    pids, slices = [], np.array_split(np.arange(len(sample)), P)
    reps = 20
    r, wfd = os.pipe()
    t2 = time.time()
    for sl in slices:
        pid = os.fork()
        if pid == 0:
            try:
                ts = time.time()
                for _ in range(reps):
                    for i in sl:
                        eng.search(sample[i])
                os.write(wfd, (f"{time.time() - ts:.6f}\n").encode())
            finally:
                os._exit(0)
        pids.append(pid)
    for pid in pids:
        os.waitpid(pid, 0)
    os.close(wfd)
    times = [float(x) for x in os.read(r, 1 << 16).decode().split()]
    os.close(r)
    multi_s = max(times) / reps if times else float("inf")
    visited = float(np.mean([res[3]["visited"] for res in results]))
    dc = float(np.mean([res[3]["distance_computations"] for res in results]))
    return dict(kind=kind, sample=desc, n_sample=len(sample), build_s=build_s, qps_1thread=len(sample) / single_s,
                qps_allcores=len(sample) / multi_s, cores=P, visited=visited, distcomp=dc, results=results, idx=idx,
                wall_s=time.time() - t0)


# ------------------------------------------------------------------------------------------------ our arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dist", default="planted", choices=["planted", "uniform"])
    ap.add_argument("--small", action="store_true", help="reduced shape for smoke runs; NOT a valid bench number")
    ap.add_argument("--queries", type=int, default=0, help="queries per step (default: the workload's own batch, 10 000)")
    ap.add_argument("--workload", default="glove100", choices=["glove100", "glove25", "readme", "deep96"],
                    help="glove100 (default) is the configuration the metric is quoted on; the others are BASELINE.json's parity-size configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="time `value` with stream-ordered calls (one batch at a time) instead of clann_search_device_async "
                         "(three batches in flight); the stream-ordered figure is always reported as value_stream_ordered")
    ap.add_argument("--shard", default="queries", choices=["queries", "clusters"],
                    help="N > 1: 'queries' = the index is replicated and every rank searches its own batch (weak scaling, no "
                         "data-path collective); 'clusters' = clusters are sharded by owner and one fixed batch is stepped "
                         "through the ranks with an all-gather per step (strong scaling; for indices beyond one GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0

    w = workload(args)
    if args.queries > 0:
        w["nq"] = args.queries
        w["name"] += f" [batch overridden: {args.queries} queries per step]"
    cfg_json = {"workload": w["name"], "distribution": args.dist, "n": w["n"], "d": w["d"], "queries_per_step": w["nq"],
                "num_tables": w["L"], "num_clusters_factor": w["factor"], "k": w["k"], "delta": w["delta"],
                "l2": "index working set (2.4 GB: Q15 rows, sketches, tables) >> 126 MB L2; no explicit flush"}

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
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * ref["n_sample"] / ref["qps_allcores"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16", "data": "synthetic",
                "config": cfg_json,
                "cpu_baseline": {"value": ref["qps_allcores"], "unit": UNIT, "cores": ref["cores"], "kind": ref["kind"],
                                 "sample": ref["sample"], "qps_1thread": ref["qps_1thread"],
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
    from clann_b200.distributed import ShardedSearcher

    data, queries, src = make_data(w, args.dist)
    nq, k, d = w["nq"], w["k"], w["d"]
    shard_clusters = world > 1 and args.shard == "clusters"
    if world > 1 and not shard_clusters:
        # weak scaling: every rank owns a full replica and its own batch of nq queries (same distribution, its own seed)
        rq = np.random.default_rng(43 + 1000 * rank)
        if args.dist == "planted":
            src = rq.integers(0, w["n"], nq)
            queries = data[src] + np.float32(0.05) * rq.standard_normal((nq, d), dtype=np.float32)
        else:
            queries = rq.standard_normal((nq, d), dtype=np.float32)
        queries /= np.linalg.norm(queries, axis=1, keepdims=True)
        queries = np.ascontiguousarray(queries, np.float32)
    cfg_json["parallelism"] = ("single GPU" if world == 1 else
                               f"clusters sharded over {world} GPUs, one batch stepped with an all-gather of query states per step"
                               if shard_clusters else
                               f"index replicated on {world} GPUs, {nq} queries per GPU per step (global batch {nq * world}), no data-path collective")

    # ---- build (untimed setup of the search benchmark; reported on its own)
    t0 = time.time()
    index = cb.init_with_config(data, cb.Config(w["L"], w["factor"], w["k"], w["delta"], "bench"))
    index.set_option("seed", 1234)
    if shard_clusters:
        index.set_option("shard_count", world)
        index.set_option("shard_rank", rank)
    index.build()
    torch.cuda.synchronize()
    build_wall = time.time() - t0
    build_ms = index.export(cl.X_BUILD_MS, 0, np.float64)
    K = index.num_clusters
    centers = index.export(cl.X_CENTERS, 0, np.uint64).copy()
    assignment = index.export(cl.X_ASSIGNMENT, 0, np.uint64).copy()
    radii = index.export(cl.X_RADII, 0, np.float32).copy()

    dev = torch.device("cuda", local_rank)
    d_q = torch.from_numpy(queries).to(dev)
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_dists = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_counts = torch.empty(nq, dtype=torch.int32, device=dev)
    searcher = ShardedSearcher(index, world, rank) if shard_clusters else ShardedSearcher(index, 1, 0)
    single = not shard_clusters  # this rank runs the whole single-GPU path on its own batch

    def step_device():
        searcher.search_device(d_q, d_ids, d_dists, d_counts)

    # batch pipelining (clann_search_device_async): consecutive steps on the internal streams (three batches in flight), each with its own outputs
    pipelined = single and not args.no_pipeline
    # 12 output sets: calls i and i + 12 land on the same internal stream for every pipeline depth (2, 3 or 4), so a set is never
    # written by two batches that could be in flight together
    NSETS = 12
    outs = [(d_ids, d_dists, d_counts)] + [(torch.empty_like(d_ids), torch.empty_like(d_dists), torch.empty_like(d_counts)) for _ in range(NSETS - 1)]
    cur_stream = torch.cuda.current_stream().cuda_stream

    def run_steps(steps):
        if not pipelined:
            for _ in range(steps):
                step_device()
            return
        for i in range(steps):
            o = outs[i % NSETS]
            if index._lib.clann_search_device_async(index.handle, d_q.data_ptr(), nq, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) != 0:
                raise RuntimeError(cl.last_error())
        if index._lib.clann_search_flush(index.handle, cur_stream) != 0:
            raise RuntimeError(cl.last_error())

    # pinned host buffers for the end-to-end arm
    h_q = torch.from_numpy(queries).pin_memory()
    h_ids = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    h_dists = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    h_counts = torch.empty(nq, dtype=torch.int32).pin_memory()

    def step_e2e():
        if single:
            # the reference-facing call: host pointers in, host pointers out (copies inside clann_search)
            st = index._lib.clann_search(index.handle, h_q.data_ptr(), nq, h_ids.data_ptr(), h_dists.data_ptr(), h_counts.data_ptr())
            if st != 0:
                raise RuntimeError(cl.last_error())
        else:
            d_q.copy_(h_q, non_blocking=True)
            searcher.search_device(d_q, d_ids, d_dists, d_counts)
            h_ids.copy_(d_ids, non_blocking=True); h_dists.copy_(d_dists, non_blocking=True); h_counts.copy_(d_counts, non_blocking=True)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        probe_ms = 0.0
        for _ in range(steps):
            fn()
            if world == 1:
                pass
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_device()
    run_steps(args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # the timed region: exactly K steps; the stream is idle at e0 (barrier + synchronize), so e0..e1 covers every kernel of
    # the K steps whichever internal stream ran it (clann_search_flush makes the current stream wait for all of them)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_ms_sum, prep_ms_sum = 0.0, 0.0
    e0.record()
    run_steps(args.steps)
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    # the same K steps one batch at a time (stream-ordered calls), for reference
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        step_device()
    f1.record()
    barrier()
    ordered_ms = f0.elapsed_time(f1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        t = torch.tensor([ordered_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ordered_ms = float(t.item())
    clocks = sampler.stop()
    launches_per_step = searcher.last_launches
    # per-kernel split from a separate pass (events inside the library); same stream, same inputs
    if single:
        for _ in range(args.steps):
            step_device()
            prof = index.search_profile()
            probe_ms_sum += prof["probe_ms"]; prep_ms_sum += prof["prep_ms"]
        probe_ms = probe_ms_sum / args.steps
        prep_ms = prep_ms_sum / args.steps
    else:
        probe_ms = prep_ms = None

    ms_per_step = total_ms / args.steps
    global_nq = nq if (world == 1 or shard_clusters) else nq * world   # queries the whole job answers per step
    value = global_nq / (ms_per_step / 1000.0)

    # end to end
    for _ in range(2):
        step_e2e()
    e2e_sync_ms = timed(step_e2e, args.steps) / args.steps
    e2e_ms, e2e_mode = e2e_sync_ms, "clann_search: one synchronous call per step (host buffers in, host buffers out)"
    if pipelined:
        # the same through clann_search_async: every step's H2D copy, search and D2H copies on its batch stream, three batches in
        # flight, each with its own pinned output buffers; clann_search_wait before the clock stops
        h_outs = [(h_ids, h_dists, h_counts)] + [(torch.empty_like(h_ids).pin_memory(), torch.empty_like(h_dists).pin_memory(),
                                                  torch.empty_like(h_counts).pin_memory()) for _ in range(NSETS - 1)]

        def run_e2e_async(steps):
            for i in range(steps):
                o = h_outs[i % NSETS]
                if index._lib.clann_search_async(index.handle, h_q.data_ptr(), nq, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()) != 0:
                    raise RuntimeError(cl.last_error())
            if index._lib.clann_search_flush(index.handle, cur_stream) != 0:
                raise RuntimeError(cl.last_error())

        run_e2e_async(3)
        if index._lib.clann_search_wait(index.handle) != 0:
            raise RuntimeError(cl.last_error())
        step_e2e()                                   # reference result of the synchronous call in h_outs[0]
        want = (h_ids.clone(), h_dists.clone(), h_counts.clone())
        run_e2e_async(NSETS)
        index._lib.clann_search_wait(index.handle)
        for o in h_outs:
            if not (torch.equal(o[0], want[0]) and torch.equal(o[1], want[1]) and torch.equal(o[2], want[2])):
                raise RuntimeError("clann_search_async returned results that differ from clann_search")
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        run_e2e_async(args.steps)
        g1.record()
        index._lib.clann_search_wait(index.handle)
        barrier()
        e2e_ms = g0.elapsed_time(g1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e_ms /= args.steps
        e2e_mode = ("clann_search_async + clann_search_wait: host buffers in and out, H2D / search / D2H of each step on its batch "
                    "stream, three batches in flight; results identical to clann_search (checked)")
    e2e_value = global_nq / (e2e_ms / 1000.0)

    # ---- correctness of what was timed: the pipelined batches return what the stream-ordered call returns
    pipe_same = None
    if pipelined:
        run_steps(3)
        torch.cuda.synchronize()
        a_ids, a_dists = outs[1][0].clone(), outs[1][1].clone()
        b_ids, b_dists = outs[2][0].clone(), outs[2][1].clone()
        step_device()
        torch.cuda.synchronize()
        pipe_same = bool(torch.equal(a_ids, d_ids) and torch.equal(b_ids, d_ids) and torch.equal(a_dists, d_dists) and torch.equal(b_dists, d_dists))
        if not pipe_same:
            raise RuntimeError("pipelined batches returned results that differ from the stream-ordered call")
    # ---- recall@k against exact fp32 neighbours (utils/mod.rs:59-95)
    step_device()
    torch.cuda.synchronize()
    ids = d_ids.cpu().numpy().view(np.uint32); dists = d_dists.cpu().numpy(); counts = d_counts.cpu().numpy()
    nchk = min(nq, 2000)
    with torch.no_grad():
        dd = torch.from_numpy(data).to(dev)
        ex = torch.empty((nchk, k), device=dev)
        chunk = max(8, min(250, int(2.5e8 // w["n"])))   # keep the similarity tile under ~1 GB
        for s in range(0, nchk, chunk):
            e = min(s + chunk, nchk)
            sim = d_q[s:e] @ dd.T
            ex[s:e] = torch.topk(sim, k, dim=1).values
        kth = (1.0 - ex[:, k - 1]).cpu().numpy()
        del dd
    hit = sum(int(np.sum(dists[i, :counts[i]] <= kth[i] + 1e-3)) for i in range(nchk))
    recall = hit / (nchk * k)

    line = None
    if rank == 0:
        ctr = index.counters(nq) if single else searcher.counters(nq)
        cand = float(ctr["candidates"].sum()); dc = float(ctr["distance_computations"].sum()); vis = float(ctr["clusters_visited"].sum())
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
            roofline = {"bound": "hbm", "kernel": "k_dense_sims + k_first_ranges + k_probe (rerank and anchors of first visits streamed, then one warp per query)", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                        "traffic": MEASURED_TRAFFIC.get(args.workload if not args.small else "small"),
                        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one k_probe launch, ncu --set full, "
                                          "k_dense_sims + k_first_ranges + k_probe, profiles/r1g_first_ranges_dense_sims_probe_full_raw.csv (glove100 planted only)",
                        "kernel_ms": probe_ms, "prep_ms": prep_ms,
                        "kernel_ms_note": "k_dense_sims (the rerank arithmetic of every query's first visit, streamed) + k_first_ranges (its "
                                          "anchors) + k_probe: the kernels between the library's events, all counted against the same "
                                          "algorithmic bytes",
                        "algorithmic_bytes_per_launch": rerank_bytes + filter_bytes,
                        "rerank_gbs": rerank_bytes / (probe_ms / 1000.0) / 1e9, "filter_gbs": filter_bytes / (probe_ms / 1000.0) / 1e9}
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is reported at N=1 only
            try:
                ref = run_reference(w, data, queries, src, centers, assignment, radii, threads=1)
                cpu = {"value": ref["qps_1thread"], "unit": UNIT, "cores": 1, "kind": ref["kind"], "sample": ref["sample"],
                       "clusters_visited_per_query": ref["visited"], "distance_computations_per_query": ref["distcomp"],
                       "index_build_s_for_sample": ref["build_s"]}
            except Exception as e:  # the baseline is reported, never required for our own number
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if shard_clusters else "weak",
            "vs_baseline": None, "dtype": "i16",
            "value_stream_ordered": global_nq / (ordered_ms / args.steps / 1000.0), "ms_per_step_stream_ordered": ordered_ms / args.steps,
            "pipeline": ("clann_search_device_async: three batches in flight on three internal streams; outputs identical to the "
                         "stream-ordered call (checked)") if pipelined else "none (stream-ordered calls)",
            "data": "synthetic", "config": cfg_json, "recall_at_k": recall, "recall_queries_checked": nchk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": global_nq * d * 4,
                    "d2h_bytes_per_step": global_nq * k * 8 + global_nq * 4,
                    "ms_per_step": e2e_ms, "call": e2e_mode,
                    "value_synchronous_call": global_nq / (e2e_sync_ms / 1000.0), "ms_per_step_synchronous_call": e2e_sync_ms},
            "gpu_launches": int(launches_per_step * args.steps * (1 if (world == 1 or shard_clusters) else world)), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "build": {"wall_s": build_wall, "gmm_ms": build_ms[0], "hash_ms": build_ms[1], "sort_ms": build_ms[2], "device_ms": build_ms[3],
                      "clusters": int(K)},
            "per_query": {"clusters_visited": vis / nq, "candidates": cand / nq, "distance_computations": dc / nq},
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
