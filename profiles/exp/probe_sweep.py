#!/usr/bin/env python
"""A/B sweep of probe-kernel launch parameters on one built index (glove-100 shape, planted queries).

    python profiles/exp/probe_sweep.py [--once] [--small] CONFIG ...

CONFIG = comma-separated key=value knobs of clann_tune (e.g. probe_warps=4,probe_ctas=1); "default" = no knob.
Without --once every configuration runs 3 warm-up + 5 timed searches and prints the mean probe-kernel time (CUDA events
recorded by the library). With --once every configuration launches exactly one search, in order — the form to run under
`ncu -k regex:k_probe --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct`.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
import clann_b200 as cb
from clann_b200 import _lib as cl


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    once = "--once" in sys.argv
    small = "--small" in sys.argv

    class A:
        pass
    a = A()
    a.small, a.workload = small, "glove100"
    w = bench.workload(a)
    data, queries, _ = bench.make_data(w, "planted")
    index = cb.init_with_config(data, cb.Config(w["L"], w["factor"], w["k"], w["delta"], "sweep"))
    index.set_option("seed", 1234)
    index.build()
    nq, k = w["nq"], w["k"]
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(queries).to(dev)
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_dists = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_counts = torch.empty(nq, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def run():
        st = index._lib.clann_search_device(index.handle, d_q.data_ptr(), nq, d_ids.data_ptr(), d_dists.data_ptr(),
                                            d_counts.data_ptr(), stream)
        if st != 0:
            raise RuntimeError(cl.last_error())

    ref = None
    seen = {}
    for cfg in args or ["default"]:
        knobs = {} if cfg == "default" else dict(kv.split("=") for kv in cfg.split(","))
        for key in seen:          # back to the default for knobs of earlier configurations
            cl.tune(key, seen[key])
        for key, v in knobs.items():
            seen.setdefault(key, DEFAULTS.get(key, 0))
            cl.tune(key, int(v))
        if once:
            run()
            torch.cuda.synchronize()
            print(f"{cfg}: launched", flush=True)
        else:
            for _ in range(3):
                run()
            ms, prep = [], []
            for _ in range(5):
                run()
                p = index.search_profile()
                ms.append(p["probe_ms"]); prep.append(p["prep_ms"])
            print(f"{cfg}: probe_ms mean {np.mean(ms):.3f} min {np.min(ms):.3f}  prep_ms {np.mean(prep):.3f}", flush=True)
        torch.cuda.synchronize()
        out = (d_ids.cpu().numpy().copy(), d_dists.cpu().numpy().view(np.uint32).copy(), d_counts.cpu().numpy().copy())
        if ref is None:
            ref = out
        else:
            same = all(np.array_equal(x, y) for x, y in zip(ref, out))
            if not same:
                print(f"{cfg}: RESULTS DIFFER from the first configuration", flush=True)


DEFAULTS = {"probe": 0, "probe_occ": 0, "probe_warps": 8, "probe_ctas": 0, "probe_nomemo": 0, "probe_cta_occ": 4,
            "probe_nosort": 0, "probe_prefetch": -2, "probe2_warps": 8, "probe2_ctas": 1, "probe2_stage_rows": 64, "l2_fetch": 64, "dense_sims": 1, "first_ranges": 2, "probe_smem_memo": 1, "order_longest_first": 0, "probe_prefetch_rows": 0, "probe2_l2": 1, "probe2_l2rows": 1, "probe2_l2idx": 1}

if __name__ == "__main__":
    main()
