"""Where the end-to-end matching time goes (host side)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from opensfm_b200 import matching
import bench

pb, feats, pairs, w = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "c4")
feats = [torch.from_numpy(f).pin_memory().numpy() for f in feats]
needed = sorted({i for p in pairs for i in p})
cfg = {"lowes_ratio": 0.8, "symmetric_matching": True}
for rep in range(3):
    t0 = time.perf_counter()
    pm = matching.PairMatcher(device=0)
    t1 = time.perf_counter()
    pm.add_many([(i, feats[i]) for i in needed])
    t2 = time.perf_counter()
    pm.submit(pairs, 0.8, True)
    t3 = time.perf_counter()
    raw = pm.fetch_raw()
    t4 = time.perf_counter()
    res = pm.match_pairs(pairs, cfg)
    t5 = time.perf_counter()
    del pm
    t6 = time.perf_counter()
    print("create %.1f add_many %.1f submit %.1f fetch %.1f  match_pairs(total, incl. dict) %.1f  destroy %.1f ms" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t6 - t5)), flush=True)
