"""Generates the committed golden fixtures.  Run in the build container:

    python tests/golden/make_golden.py

* match_golden.npz — outputs of the *reference matcher* (cv2 4.13.0 through the verbatim
  opensfm/matching.py:723-777 code in oracle/match_oracle.py) on seeded inputs that
  opensfm_b200.synthetic regenerates from the stored seeds.
* ba_golden.npz — outputs of the BA oracle (restated Ceres path; Ceres itself cannot run in
  this image) on a seeded cube scene: costs, iteration count, solution arrays.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from opensfm_b200 import synthetic as syn  # noqa: E402
from oracle import ba_lm, match_oracle as mo  # noqa: E402


def match_cases():
    """(name, f1, f2, mask) regenerated identically by tests/test_golden.py."""
    a = syn.hahog_like_descriptors(900, 101)
    b = syn.hahog_like_descriptors(1100, 102)
    b[:400] = np.clip(a[:400] + np.random.RandomState(103).randint(-6, 7, (400, 128)), 0, 255).astype(np.float32)
    b[1000:1010] = b[0:10]
    yield "hahog", a, b, None
    mask = np.random.RandomState(104).rand(900, 1100) < 0.04
    yield "hahog_mask", a, b, mask
    sc = syn.cube_scene(4, 1200, 1.0)
    yield "cube01", sc.features_of_shot(0)[0], sc.features_of_shot(1)[0], None
    u1 = syn.binary_descriptors(600, 105)
    u2 = syn.binary_descriptors(800, 106)
    u2[:250] = u1[:250]
    u2[:250, :2] ^= 9
    yield "akaze", u1, u2, None
    rng = np.random.RandomState(107)
    g1 = rng.rand(300, 128).astype(np.float32)
    g2 = rng.rand(400, 128).astype(np.float32)
    g2[:150] = g1[:150] + rng.normal(0, 0.02, (150, 128)).astype(np.float32)
    yield "float", g1, g2, None


def ba_case():
    sc = syn.cube_scene(8, 400, 1.0, with_descriptors=False)
    return syn.scene_to_problem(sc)


if __name__ == "__main__":
    import cv2

    cfg = {"lowes_ratio": 0.8}
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, f1, f2, mask in match_cases():
        out[name + "_oneway"] = np.array(mo.match_brute_force(f1, f2, cfg, mask), dtype=np.int32).reshape(-1, 2)
        out[name + "_sym"] = np.array(sorted(mo.match_brute_force_symmetric(f1, f2, cfg, mask)), dtype=np.int32).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "match_golden.npz"), **out)
    res = ba_lm.solve(ba_case())
    np.savez_compressed(os.path.join(HERE, "ba_golden.npz"), initial_cost=res["initial_cost"],
                        final_cost=res["final_cost"], iterations=res["iterations"], points=res["points"],
                        inst=res["inst"], cam_params=res["cam_params"],
                        reprojection_errors=res["reprojection_errors"])
    print("wrote", sorted(out), "and ba_golden.npz")
