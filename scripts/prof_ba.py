"""A bounded BA run (for ncu): the C4 scene, 2 LM iterations."""
import sys
sys.path.insert(0, ".")
from opensfm_b200 import bundle, synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
cfg = {"c4": (500, 200000, 10), "mid": (100, 20000, 10), "c2": (50, 5000, None)}[name]
sc = syn.cube_scene(cfg[0], cfg[1], 1.0, with_descriptors=False, max_obs_per_point=cfg[2])
pb = syn.scene_to_problem(sc)
pb.max_iterations = 2
res = bundle.solve(pb)
print(res["summary"])
