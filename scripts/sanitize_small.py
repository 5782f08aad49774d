"""Small invocations of every hot kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from opensfm_b200 import bundle, matching, synthetic as syn  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
cfg = {"lowes_ratio": 0.8}
if what in ("all", "ba"):
    sc = syn.cube_scene(12, 1500, 1.0, with_descriptors=False, max_obs_per_point=8)
    r = bundle.solve(syn.scene_to_problem(sc))
    print("ba", r["summary"]["termination"], r["summary"]["iterations"], r["summary"]["final_cost"])
if what in ("all", "match"):
    sc = syn.cube_scene(3, 700, 1.0)
    f = [sc.features_of_shot(i)[0] for i in range(3)]
    pm = matching.PairMatcher()
    pm.add_many([(i, x.astype(np.uint8)) for i, x in enumerate(f)], uint8_is_l2=True)
    out = pm.match_pairs([(0, 1), (1, 2)], cfg)
    print("l2", pm.last_kernel(), [len(v) for v in out.values()])
    b = [syn.binary_descriptors(500, 10 + i, 61) for i in range(2)]
    print("h8", len(matching.match_brute_force_symmetric(b[0], b[1], cfg)))
    g = np.random.RandomState(0).rand(300, 128).astype(np.float32)
    print("f32", len(matching.match_brute_force(g, g[::-1].copy(), cfg)))
