"""Solve one scene with the persistent Schur kernel and with the CTA-per-segment kernel it replaces; compare."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = (
    "import sys; sys.path.insert(0, %r)\n"
    "import numpy as np\n"
    "from opensfm_b200 import bundle, synthetic as syn\n"
    "sc = syn.cube_scene(30, 4000, 1.0, with_descriptors=False, max_obs_per_point=8)\n"
    "r = bundle.solve(syn.scene_to_problem(sc))\n"
    "print(r['summary']['final_cost'], r['summary']['iterations'])\n"
    "np.save(sys.argv[1], r['points'])\n"
) % ROOT
import numpy as np
out = {}
for name, env in (("pipe", {}), ("mma", {"OSFM_BA_SCHUR_PIPE": "0"})):
    p = subprocess.run([sys.executable, "-c", code, "/tmp/chk_%s.npy" % name], env=dict(os.environ, **env), timeout=120,
                       capture_output=True, text=True)
    print(name, p.returncode, p.stdout.strip(), p.stderr.strip()[-400:])
    if p.returncode == 0:
        out[name] = np.load("/tmp/chk_%s.npy" % name)
if len(out) == 2:
    print("max point difference", np.abs(out["pipe"] - out["mma"]).max())
    sys.exit(0)
sys.exit(1)
