"""Per-phase trace of one BA solve on the BASELINE configs[3] workload (OSFM_BA_TRACE=1 prints to stderr)."""
import os
import sys

os.environ["OSFM_BA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from opensfm_b200 import bundle  # noqa: E402

pb, feats, pairs, w = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "c4")
bundle.solve(pb)
r = bundle.solve(pb)
print({k: v for k, v in r["summary"].items() if k.startswith("time") or k.endswith("launches") or k in ("iterations", "pcg_iterations")})
