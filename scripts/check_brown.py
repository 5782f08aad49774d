import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import numpy as np
import test_ba_gpu as t
from opensfm_b200 import bundle, ba_problem as bp
pb = t._mixed_problem(3, only=getattr(bp, sys.argv[1]))
r = bundle.solve(pb)
s = r["summary"]
print(sys.argv[1], s["termination"], s["iterations"], s["initial_cost"], s["final_cost"], s["message"])
''' % (ROOT, ROOT)
for ptype in ("BROWN", "FISHEYE"):
    for name, env in (("default", {}), ("generic_lin", {"OSFM_BA_LIN_SPECIAL": "0"}), ("generic_lin+point_schur", {"OSFM_BA_LIN_SPECIAL": "0", "OSFM_BA_SEGMENT_SCHUR": "0"}),
                      ("special_lin+point_schur", {"OSFM_BA_SEGMENT_SCHUR": "0"}), ("special+simt_seg", {"OSFM_BA_SCHUR_MMA": "0"})):
        p = subprocess.run([sys.executable, "-c", code, ptype], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        print(name, "|", p.stdout.strip(), p.stderr.strip()[-300:])
