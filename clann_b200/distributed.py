"""Multi-GPU search: one process per GPU, clusters sharded by owner, queries replicated.

Two protocols, both implemented in libclann_b200.so; this module only drives them from Python:

* ClusterShardedSearcher — the measured mode (clann_search_sharded and its multi-batch / streaming forms):
  queries are routed to the owner of their nearest cluster, which runs the reference's loop (src/core/index.rs:331-432) for as long
  as the walk stays in its own clusters; one all-reduce(min) of a per-query bound; a pruned second round on every rank; one
  all-gather of the top-k lists and a k-way merge. The collectives are issued inside the library (NCCL, or two callbacks: see
  InProcessTransport). torch.distributed only carries the 128-byte NCCL id at start-up.
* ShardedSearcher — the exact stepping protocol: every rank holds the full visiting order; in each step a rank advances every
  unfinished query through the consecutive clusters it owns and stops at the first foreign one; one all-gather of the per-query
  state (heap + position, 32 + 8k bytes) hands each query to the owner of its next cluster. Bit-identical to the single-GPU search,
  one collective per hand-over; the control loop below is backend-agnostic so it can be exercised on CPU with gloo (tests/).

The small numpy helpers (order_bits, pack_bound, merge_topk) restate the device's exchange formats for the CPU tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable

import numpy as np

from . import _lib
from .api import ClusteredIndex, _check


def run_stepping(step: Callable[[], None], exchange_and_merge: Callable[[], int], max_steps: int) -> int:
    """Advance-until-foreign / exchange loop. `exchange_and_merge` returns the number of queries still active anywhere.
    Returns the number of steps taken. max_steps bounds it (a query changes owner at most K-1 times)."""
    steps = 0
    while True:
        step()
        steps += 1
        active = exchange_and_merge()
        if active == 0:
            return steps
        if steps > max_steps:
            raise RuntimeError(f"sharded search did not converge after {steps} steps ({active} queries active)")


class _DeviceBytes:
    """Zero-copy torch view of device memory owned by the library (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class ShardedSearcher:
    def __init__(self, index: ClusteredIndex, world: int = 1, rank: int = 0):
        self.index, self.world, self.rank = index, world, rank
        self.lib = _lib.load()
        self.last_launches = 0
        self.last_steps = 0
        self._gather = None

    def search_device(self, d_queries, d_ids, d_dists, d_counts) -> None:
        """d_* are torch CUDA tensors (queries [nq,d] f32; ids [nq,k] i32; dists [nq,k] f32; counts [nq] i32)."""
        import torch
        h = self.index.handle
        nq = d_queries.shape[0]
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.world == 1:
            _check(self.lib.clann_search_device(h, d_queries.data_ptr(), nq, d_ids.data_ptr(), d_dists.data_ptr(),
                                                d_counts.data_ptr(), stream))
            launches = C.c_uint32(0)
            _check(self.lib.clann_last_search_profile(h, None, C.byref(launches)))  # kernels the library launched for this call
            self.last_launches, self.last_steps = int(launches.value), 1
            return
        import torch.distributed as dist
        _check(self.lib.clann_search_begin(h, d_queries.data_ptr(), nq, stream))
        sb = int(self.lib.clann_state_bytes(h))
        local = torch.as_tensor(_DeviceBytes(self.lib.clann_state_ptr(h), nq * sb), device=d_queries.device)
        if self._gather is None or self._gather.numel() != self.world * nq * sb:
            self._gather = torch.empty(self.world * nq * sb, dtype=torch.uint8, device=d_queries.device)
        launches = [7]

        def step():
            _check(self.lib.clann_search_step(h, stream))
            launches[0] += 1

        def exchange():
            dist.all_gather_into_tensor(self._gather, local)
            active = C.c_uint64(0)
            _check(self.lib.clann_search_merge(h, self._gather.data_ptr(), self.world, C.byref(active), stream))
            launches[0] += 1
            return int(active.value)

        self.last_steps = run_stepping(step, exchange, max_steps=self.index.num_clusters + 1)
        _check(self.lib.clann_search_end(h, d_ids.data_ptr(), d_dists.data_ptr(), d_counts.data_ptr(), stream))
        self.last_launches = launches[0] + 1

    def counters(self, nq: int):
        return self.index.counters(nq)


# ------------------------------------------------------------------------------------------------ cluster-sharded search (fast mode)

INF_BITS = 0xFF800000          # order bits of +inf (clann_b200/csrc/common.cuh float_order_bits)
NOTHING = (INF_BITS << 32) | 0xFFFFFFFF


def order_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> uint32 preserving order (float_order_bits in common.cuh)."""
    b = np.asarray(x, np.float32).view(np.uint32)
    return np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)


def pack_bound(bound: np.ndarray, consumed: np.ndarray) -> np.ndarray:
    """(bound on the k-th distance, clusters consumed in round one) -> the u64 word the ranks min-reduce (k_shard_pack_bounds):
    the smaller bound wins, at equal bounds the rank that consumed more."""
    return (order_bits(bound).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(consumed, np.uint64))


def u64_min_key(x):
    """torch has no unsigned 64-bit min: flip the top bit and compare as int64 (order preserving)."""
    import torch
    return torch.bitwise_xor(x.view(torch.int64), torch.tensor(-2 ** 63, dtype=torch.int64, device=x.device))


def merge_topk(lists: np.ndarray, k: int) -> np.ndarray:
    """k-way merge of per-rank candidate keys ((order_bits(dist) << 32) | id, ~0 = empty), as k_shard_final_merge does:
    lists [world, nq, k] -> [nq, k] ascending."""
    world, nq, kk = lists.shape
    flat = np.transpose(lists, (1, 0, 2)).reshape(nq, world * kk)
    return np.sort(flat, axis=1)[:, :k]


class ClusterShardedSearcher:
    """bench.py / user-facing driver of clann_search_sharded with the NCCL transport: the 128-byte NCCL unique id is created on
    rank 0 and broadcast with torch.distributed (any channel would do), then every rank creates its communicator inside the
    library. All ranks must build the index with shard_count = world, shard_rank = rank over the same data."""

    def __init__(self, index: ClusteredIndex, world: int, rank: int):
        import torch
        import torch.distributed as dist
        self.index, self.world, self.rank = index, world, rank
        self.lib = _lib.load()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            _check(self.lib.clann_comm_unique_id(buf, 128))
            uid = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device())
        uid_dev = uid.to(dev)
        dist.broadcast(uid_dev, src=0)
        host = uid_dev.cpu().numpy()
        _check(self.lib.clann_comm_init(index.handle, rank, world, host.ctypes.data))

    def search_device(self, d_queries, d_ids, d_dists, d_counts) -> None:
        import torch
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _check(self.lib.clann_search_sharded(self.index.handle, d_queries.data_ptr(), d_queries.shape[0], d_ids.data_ptr(),
                                             d_dists.data_ptr(), d_counts.data_ptr(), stream))

    def search_device_pair(self, q_a, q_b, out_a, out_b) -> None:
        """Two whole batches in flight (clann_search_sharded_pair); out_* = (ids, dists, counts) tensors."""
        import torch
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _check(self.lib.clann_search_sharded_pair(self.index.handle, q_a.data_ptr(), q_b.data_ptr(), q_a.shape[0], out_a[0].data_ptr(),
                                                  out_a[1].data_ptr(), out_a[2].data_ptr(), out_b[0].data_ptr(), out_b[1].data_ptr(),
                                                  out_b[2].data_ptr(), stream))

    def search_device_multi(self, queries, outs) -> None:
        """1 to 4 whole batches in flight (clann_search_sharded_multi): queries = list of [nq, d] tensors, outs = list of
        (ids, dists, counts) tensor triples."""
        import torch
        nb = len(queries)
        arr = lambda ptrs: (C.c_void_p * nb)(*ptrs)  # noqa: E731
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _check(self.lib.clann_search_sharded_multi(self.index.handle, nb, arr([q.data_ptr() for q in queries]), queries[0].shape[0],
                                                   arr([o[0].data_ptr() for o in outs]), arr([o[1].data_ptr() for o in outs]),
                                                   arr([o[2].data_ptr() for o in outs]), stream))

    def submit(self, d_queries, out) -> None:
        """Streaming form (clann_search_sharded_submit): one more batch into the software pipeline; out = (ids, dists, counts).
        The batch's buffers must stay untouched until three more submits or flush(); every rank makes the same calls."""
        import torch
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _check(self.lib.clann_search_sharded_submit(self.index.handle, d_queries.data_ptr(), d_queries.shape[0], out[0].data_ptr(),
                                                    out[1].data_ptr(), out[2].data_ptr(), stream))

    def flush(self) -> None:
        import torch
        _check(self.lib.clann_search_sharded_flush(self.index.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        ms = (C.c_float * 6)()
        _check(self.lib.clann_shard_stats(self.index.handle, C.byref(a), C.byref(b), ms))
        self.phase_ms = dict(zip(("route_scoring", "route_exchange", "round_one", "bound_exchange", "round_two", "merge"), [float(x) for x in ms]))
        return int(a.value), int(b.value)


class InProcessTransport:
    """Collectives for `world` indices living in ONE process (one thread per rank, any number of GPUs including one): the
    clann_set_collectives callbacks copy through a shared staging tensor between two thread barriers. Kernels of different ranks
    never wait on each other on the device — only the host threads meet — so this is safe on a single GPU. Used by the tests."""

    def __init__(self, world: int, device):
        import threading
        import torch
        self.world, self.device = world, device
        self.barrier = threading.Barrier(world)
        self.stage = None
        self.lock = threading.Lock()
        self.torch = torch
        self._keep = []

    def _ensure(self, nbytes):
        with self.lock:
            if self.stage is None or self.stage.shape[1] < nbytes:
                self.stage = self.torch.empty((self.world, nbytes), dtype=self.torch.uint8, device=self.device)

    def _view(self, ptr, nbytes):
        return self.torch.as_tensor(_DeviceBytes(ptr, nbytes), device=self.device)

    def callbacks(self, rank: int):
        t = self.torch

        def allgather(ctx, send, recv, nbytes, stream):
            try:
                t.cuda.synchronize()
                self.barrier.wait()
                if rank == 0:
                    self._ensure(nbytes)
                self.barrier.wait()
                self.stage[rank, :nbytes].copy_(self._view(send, nbytes))
                t.cuda.synchronize()
                self.barrier.wait()
                self._view(recv, nbytes * self.world).copy_(self.stage[:, :nbytes].reshape(-1))
                t.cuda.synchronize()
                self.barrier.wait()
                return 0
            except Exception:  # a broken barrier / CUDA error must surface as a failed collective, not a hang
                self.barrier.abort()
                return 1

        def allreduce_min(ctx, buf, count, stream):
            try:
                nbytes = count * 8
                t.cuda.synchronize()
                self.barrier.wait()
                if rank == 0:
                    self._ensure(nbytes)
                self.barrier.wait()
                self.stage[rank, :nbytes].copy_(self._view(buf, nbytes))
                t.cuda.synchronize()
                self.barrier.wait()
                keys = u64_min_key(self.stage[:, :nbytes].contiguous().view(t.int64).reshape(self.world, count))
                best = t.bitwise_xor(keys.min(dim=0).values, t.tensor(-2 ** 63, dtype=t.int64, device=self.device))
                self._view(buf, nbytes).copy_(best.view(t.uint8))
                t.cuda.synchronize()
                self.barrier.wait()
                return 0
            except Exception:
                self.barrier.abort()
                return 1

        ag, ar = _lib.ALLGATHER_FN(allgather), _lib.ALLREDUCE_MIN_FN(allreduce_min)
        self._keep += [ag, ar]  # ctypes callbacks must outlive the calls
        return ag, ar
