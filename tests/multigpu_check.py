"""Multi-GPU check of the cluster-sharded build and search over NCCL (run with torchrun on >= 2 GPUs; not a pytest file: the
driver's GPU tests run on one GPU, where tests/test_gpu_more.py covers the same protocol with an in-process transport).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py

Checks, on every rank: the sharded clustering (each rank its share of the rows, all-reduce of the arg-max key) equals the
single-GPU clustering bit for bit; clann_search_sharded returns the same merged result on every rank, identical to the
single-GPU search for queries whose walk stays in one cluster, with a superset of visits and recall >= single-GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402
import clann_b200 as cb  # noqa: E402
from clann_b200 import _lib as cl  # noqa: E402
from clann_b200.distributed import ClusterShardedSearcher  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    data = util.planted(200_000, 96, 21)
    q = np.concatenate([util.planted_queries(data, 2000, 123), util.uniform_sphere(16, 96, 124)])
    nq, k = len(q), 10
    conf = cb.Config(84, 0.4, k, 0.9, "multigpu")
    full = cb.init_with_config(data, conf)
    full.set_option("seed", 5)
    full.build()
    ids0, d0, c0 = full.search_batch(q)
    ctr0 = full.counters(nq)
    cen0, asg0, rad0 = (full.export(cl.X_CENTERS, 0, np.uint64).copy(), full.export(cl.X_ASSIGNMENT, 0, np.uint64).copy(),
                        full.export(cl.X_RADII, 0, np.float32).copy())
    ix = cb.init_with_config(data, conf)
    ix.set_option("seed", 5)
    ix.set_option("shard_count", world)
    ix.set_option("shard_rank", rank)
    searcher = ClusterShardedSearcher(ix, world, rank)   # communicator before the build: sharded clustering
    ix.build()
    assert np.array_equal(ix.export(cl.X_CENTERS, 0, np.uint64), cen0), "sharded clustering: centres differ"
    assert np.array_equal(ix.export(cl.X_ASSIGNMENT, 0, np.uint64), asg0), "sharded clustering: assignment differs"
    assert np.array_equal(ix.export(cl.X_RADII, 0, np.float32).view(np.uint32), rad0.view(np.uint32)), "sharded clustering: radii differ"
    dq = torch.from_numpy(q).to(dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cc = torch.empty(nq, dtype=torch.int32, device=dev)
    for _ in range(2):
        searcher.search_device(dq, ids, dd, cc)
        torch.cuda.synchronize()
    # two batches in flight return what two separate calls return
    q2 = np.ascontiguousarray(q[::-1])
    dq2 = torch.from_numpy(q2).to(dev)
    outs_b = (torch.empty_like(ids), torch.empty_like(dd), torch.empty_like(cc))
    want_b = (torch.empty_like(ids), torch.empty_like(dd), torch.empty_like(cc))
    searcher.search_device(dq2, *want_b)
    pair_a = (torch.empty_like(ids), torch.empty_like(dd), torch.empty_like(cc))
    searcher.search_device_pair(dq, dq2, pair_a, outs_b)
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(pair_a, (ids, dd, cc))), "pair: first batch differs"
    assert all(torch.equal(x, y) for x, y in zip(outs_b, want_b)), "pair: second batch differs"
    # streaming form: seven batches through the staggered pipeline; every batch gets the results of the blocking call
    seq = [dq, dq2, dq, dq, dq2, dq2, dq]
    so = [(torch.empty_like(ids), torch.empty_like(dd), torch.empty_like(cc)) for _ in seq]
    for x, o in zip(seq, so):
        searcher.submit(x, o)
    searcher.flush()
    torch.cuda.synchronize()
    for x, o in zip(seq, so):
        want = (ids, dd, cc) if x is dq else want_b
        assert all(torch.equal(a, b) for a, b in zip(o, want)), "stream: a batch differs from the blocking call"
    searcher.search_device(dq, ids, dd, cc)   # counters below are those of the single call
    torch.cuda.synchronize()
    ids, dd, cc = ids.cpu().numpy().view(np.uint32), dd.cpu().numpy(), cc.cpu().numpy().view(np.uint32)
    # every rank holds the same merged result
    gathered = [torch.empty_like(torch.from_numpy(ids.view(np.int32))).to(dev) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(ids.view(np.int32)).to(dev))
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree on the merged ids"
    ctr = ix.counters(nq)
    vis = torch.from_numpy(ctr["clusters_visited"].astype(np.int64)).to(dev)
    dist.all_reduce(vis)
    vis = vis.cpu().numpy()
    one = ctr0["clusters_visited"] == 1
    assert one.sum() > 1000
    assert np.array_equal(ids[one], ids0[one]) and np.array_equal(dd[one].view(np.uint32), d0[one].view(np.uint32)) and np.array_equal(cc[one], c0[one])
    assert np.all(vis >= ctr0["clusters_visited"]), "a visit of the single-GPU search was not made"
    sample = np.arange(0, nq, 5)
    rec_s = util.recall_at_k(data, q[sample], dd[sample], cc[sample], k)
    rec_1 = util.recall_at_k(data, q[sample], d0[sample], c0[sample], k)
    assert rec_s >= rec_1 >= 0.9, (rec_s, rec_1)
    routed, still_open = searcher.stats()
    if rank == 0:
        print(f"multigpu_check ok: world {world}, clustering identical, {int(one.sum())} one-cluster queries identical, recall sharded {rec_s:.4f} "
              f">= single {rec_1:.4f}, routed to rank 0 {routed}, open after round one {still_open}, phases {searcher.phase_ms}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
