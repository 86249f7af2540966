"""Multi-GPU search: one process per GPU, clusters sharded by owner, queries replicated.

The reference's per-query loop (src/core/index.rs:331-432) is sequential over clusters: each visit needs the running top-k
heap (the prune test :342-361 and max_sim :382-389). Sharding keeps that exact: every rank holds the full visiting order;
in each step a rank advances every unfinished query through the consecutive clusters it owns and stops at the first
foreign one; one all-gather of the per-query state (heap + position, 32 + 8k bytes) then hands each query to the owner
of its next cluster. Results are therefore identical to the single-GPU search, and every (query, cluster) visit is done
exactly once, by the GPU that holds the cluster.

torch.distributed (NCCL over NVLink) is the plumbing for the one collective the path has; the kernels are in
libclann_b200.so. The control loop below is backend-agnostic so it can be exercised on CPU with gloo (tests/).
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
