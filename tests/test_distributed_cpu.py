"""World-size-2 gloo test of the multi-GPU control flow (clann_b200/distributed.py): advance-until-foreign steps with an
all-gather + merge between them must give exactly the result of the sequential single-process walk, and every
(query, cluster) visit must be executed exactly once, by the owner of the cluster."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clann_b200.distributed import run_stepping


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeEngine:
    """Numpy stand-in for the probe kernel: state per query = (pos, done, acc); visiting cluster c adds weight[q, c];
    a query stops after `budget[q]` visits (the early exit)."""

    def __init__(self, rank, owner, order, weight, budget):
        self.rank, self.owner, self.order, self.weight, self.budget = rank, owner, order, weight, budget
        nq = order.shape[0]
        self.state = np.zeros((nq, 3), np.int64)
        self.visits = []

    def step(self):
        K = self.order.shape[1]
        for q in range(self.state.shape[0]):
            pos, done, acc = self.state[q]
            if done:
                continue
            while pos < K:
                if pos >= self.budget[q]:
                    done = 1
                    break
                c = self.order[q, pos]
                if self.owner[c] != self.rank:
                    break
                acc += self.weight[q, c]
                self.visits.append((q, int(c)))
                pos += 1
            if pos >= K:
                done = 1
            self.state[q] = (pos, done, acc)

    def exchange_and_merge(self):
        mine = torch.from_numpy(self.state.copy())
        gathered = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(gathered, mine)
        allst = torch.stack(gathered).numpy()  # [world, nq, 3]
        # adopt the state of the rank that advanced the query furthest (done first, then position)
        key = allst[:, :, 1] * (1 << 40) + allst[:, :, 0]
        best = key.argmax(axis=0)
        self.state = allst[best, np.arange(allst.shape[1])]
        return int((self.state[:, 1] == 0).sum())


def _worker(rank, world, port, owner, order, weight, budget, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = FakeEngine(rank, owner, order, weight, budget)
    steps = run_stepping(eng.step, eng.exchange_and_merge, max_steps=order.shape[1] + 1)
    out[rank] = (eng.state.copy(), list(eng.visits), steps)
    dist.barrier()
    dist.destroy_process_group()


def test_stepping_matches_sequential_walk():
    rng = np.random.default_rng(0)
    nq, K, world = 37, 11, 2
    owner = rng.integers(0, world, K)
    order = np.stack([rng.permutation(K) for _ in range(nq)])
    weight = rng.integers(1, 100, (nq, K))
    budget = rng.integers(1, K + 2, nq)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), owner, order, weight, budget, out), nprocs=world, join=True)
        res = dict(out)
    # sequential reference
    exp_acc = np.array([weight[q, order[q, : min(budget[q], K)]].sum() for q in range(nq)])
    exp_pos = np.minimum(budget, K)
    for r in range(world):
        st, visits, steps = res[r]
        assert np.array_equal(st[:, 2], exp_acc) and np.array_equal(st[:, 0], exp_pos) and st[:, 1].all()
        assert all(owner[c] == r for _, c in visits)
    all_visits = res[0][1] + res[1][1]
    assert len(all_visits) == len(set(all_visits)) == int(exp_pos.sum())  # every visit exactly once
    assert res[0][2] == res[1][2] <= K + 1


def test_stepping_reports_non_convergence():
    with pytest.raises(RuntimeError, match="did not converge"):
        run_stepping(lambda: None, lambda: 1, max_steps=3)


def _sharded_protocol_worker(rank, world, port, out):
    """Host side of clann_search_sharded's exchanges under gloo: the packed (bound, consumed) words are min-reduced through the
    order-preserving int64 view, the per-rank candidate lists are all-gathered and k-way merged."""
    import torch
    import torch.distributed as dist
    from clann_b200 import distributed as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(100)           # same draws on every rank
    nq, k = 64, 5
    owner = rng.integers(0, world, nq)         # which rank advanced each query in round one
    bound = rng.random(nq).astype(np.float32)
    bound[::7] = -1.0                          # finished queries
    bound[3::11] = np.inf                      # heap not full yet
    consumed = rng.integers(1, 9, nq)
    mine = owner == rank
    packed = np.full(nq, D.NOTHING, np.uint64)
    packed[mine] = D.pack_bound(bound[mine], consumed[mine])
    t = torch.from_numpy(packed.view(np.int64).copy())
    key = D.u64_min_key(t)
    dist.all_reduce(key, op=dist.ReduceOp.MIN)
    agreed = torch.bitwise_xor(key, torch.tensor(-2 ** 63, dtype=torch.int64)).numpy().view(np.uint64)
    want = D.pack_bound(bound, consumed)
    assert np.array_equal(agreed, want)
    # open queries = non-negative bound; decode what round two reads
    open_q = (agreed >> np.uint64(32)).astype(np.uint32) >= 0x80000000
    assert np.array_equal(open_q, bound >= 0)
    assert np.array_equal((np.uint64(0xFFFFFFFF) - (agreed & np.uint64(0xFFFFFFFF))).astype(np.int64), consumed)
    # merge: every rank contributes its own candidates of every query
    allc = (D.order_bits(rng.random((world, nq, k)).astype(np.float32)).astype(np.uint64) << np.uint64(32)) | \
        rng.integers(0, 1 << 20, (world, nq, k)).astype(np.uint64)
    allc.sort(axis=2)
    allc[:, ::5, 3:] = np.uint64(0xFFFFFFFFFFFFFFFF)   # short lists
    local = torch.from_numpy(allc[rank].view(np.int64).copy())
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    lists = np.stack([g.numpy().view(np.uint64) for g in gathered])
    merged = D.merge_topk(lists, k)
    assert np.array_equal(merged, np.sort(allc.transpose(1, 0, 2).reshape(nq, -1), axis=1)[:, :k])
    out.put((rank, True))
    dist.destroy_process_group()


def test_sharded_protocol_exchanges_under_gloo():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_sharded_protocol_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    got = sorted(out.get(timeout=5) for _ in range(world))
    assert got == [(0, True), (1, True)]
