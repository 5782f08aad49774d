"""Host-side cost of one bundle.solve() call (run with OSFM_BA_TRACE=1 for the C++ phases)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from opensfm_b200 import bundle
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
pb, feats, pairs, w = bench.build_workload(wl)
for rep in range(3):
    t0 = time.perf_counter(); pb.validate(); t1 = time.perf_counter()
    r = bundle.solve(pb); t2 = time.perf_counter()
    s = r["summary"]
    print("rep %d validate %.1f ms solve %.1f ms  run %.1f ms device %.1f ms iters %d" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * s["time_run_s"], s["time_device_ms"], s["iterations"]), flush=True)
