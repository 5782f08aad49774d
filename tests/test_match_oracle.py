"""Pins the numpy restatement of cv2.BFMatcher.knnMatch against live cv2 (the real reference)."""
import numpy as np
import pytest

from oracle import match_oracle as mo
from opensfm_b200 import synthetic as syn

cv2 = pytest.importorskip("cv2")
CFG = {"lowes_ratio": 0.8}


def _related(n1, n2, seed):
    a = syn.hahog_like_descriptors(n1, seed)
    b = syn.hahog_like_descriptors(n2, seed + 1)
    k = min(n1, n2) // 2
    b[:k] = np.clip(a[:k] + np.random.RandomState(seed + 2).randint(-6, 7, (k, 128)), 0, 255).astype(np.float32)
    return a, b


def test_l2_restated_equals_cv2():
    a, b = _related(700, 900, 1)
    b[800:805] = b[0:5]
    for ratio in (0.6, 0.8, 1.0, 1.2):
        cfg = {"lowes_ratio": ratio}
        assert mo.match_brute_force_numpy(a, b, cfg) == mo.match_brute_force(a, b, cfg)


def test_ties_go_to_lowest_train_index():
    a = syn.hahog_like_descriptors(50, 3)
    b = np.concatenate([a, a], axis=0)  # every query has two exact zero-distance candidates
    m = cv2.DescriptorMatcher_create("BruteForce")
    m.add([b])
    res = m.knnMatch(a, k=2)
    assert all(r[0].trainIdx == i and r[1].trainIdx == i + 50 for i, r in enumerate(res))
    i1, d1, i2, d2 = mo.knn2_numpy(a, b)
    assert (i1 == np.arange(50)).all() and (i2 == np.arange(50) + 50).all()


def test_mask_and_short_rows():
    a, b = _related(300, 350, 5)
    mask = np.random.RandomState(0).rand(300, 350) < 0.03
    mask[0] = False
    mask[1] = False
    mask[1, 7] = True
    assert mo.match_brute_force_numpy(a, b, CFG, mask) == mo.match_brute_force(a, b, CFG, mask)
    assert sorted(mo.match_brute_force_symmetric_numpy(a, b, CFG, mask)) == sorted(
        mo.match_brute_force_symmetric(a, b, CFG, mask))


def test_hamming_restated_equals_cv2():
    u1 = syn.binary_descriptors(400, 1)
    u2 = syn.binary_descriptors(500, 2)
    u2[:150] = u1[:150]
    u2[:150, :2] ^= 3
    assert mo.match_brute_force_numpy(u1, u2, CFG) == mo.match_brute_force(u1, u2, CFG)


def test_distance_is_sqrt_of_float32_sum():
    a, b = _related(40, 60, 9)
    m = cv2.DescriptorMatcher_create("BruteForce")
    m.add([b])
    res = m.knnMatch(a, k=1)
    d = mo.distance_matrix(a, b)
    for i, r in enumerate(res):
        assert np.float32(r[0].distance) == d[i, r[0].trainIdx]


def test_cube_scene_descriptors():
    sc = syn.cube_scene(3, 800, 1.0)
    f0, f1 = sc.features_of_shot(0)[0], sc.features_of_shot(1)[0]
    assert f0.dtype == np.float32 and f0.shape[1] == 128
    assert mo.match_brute_force_numpy(f0, f1, CFG) == mo.match_brute_force(f0, f1, CFG)


@pytest.mark.parametrize("dim", [3, 16, 20, 61, 64, 100, 128, 130, 256])
def test_cv2_float_sum_order(dim):
    """The restated summation order of cv2's float L2 (4 lanes x 4 accumulators, mul then add, tree reduce, scalar
    tail) reproduces cv2's distances BIT FOR BIT on arbitrary floats -- this is what makes general-float
    matching exact by contract rather than up to near-ties."""
    rng = np.random.RandomState(dim)
    a = rng.randn(37, dim).astype(np.float32) * 3
    b = rng.randn(53, dim).astype(np.float32) * 3
    m = cv2.DescriptorMatcher_create("BruteForce")
    m.add([b])
    res = m.knnMatch(a, k=53)
    D = np.zeros((37, 53), np.float32)
    for q, lst in enumerate(res):
        for dm in lst:
            D[q, dm.trainIdx] = dm.distance
    assert np.array_equal(D, mo.distance_matrix(a, b))


def test_general_float_matches_cv2_exactly_with_near_ties():
    rng = np.random.RandomState(5)
    a = rng.rand(1500, 128).astype(np.float32)
    b = rng.rand(2000, 128).astype(np.float32)
    b[:700] = a[:700] + rng.normal(0, 0.02, (700, 128)).astype(np.float32)
    b[1500:1600] = b[:100] + np.float32(1e-7)   # near-duplicates: distances differ in the last bits only
    for ratio in (0.8, 1.0):
        cfg = {"lowes_ratio": ratio}
        assert mo.match_brute_force_numpy(a, b, cfg) == mo.match_brute_force(a, b, cfg)


def words_known_answer_case():
    """The inputs of the reference's own known-answer test for the WORDS matcher (opensfm/test/test_matching.py:23-70):
    f1 random normal rows of the unit-norm matrix, f2 = f1 + noise / 500, the `bow_words_to_match` = 50 closest visual
    words of every feature in the reference's 10000-word vocabulary (tests/golden/words_golden.npz, made from that file
    by tests/golden/make_words_golden.py).  Expected answer there: every feature matches its own copy,
    `len(matches) == nfeatures` and `i == j` for all of them."""
    import importlib.util
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_words_golden", os.path.join(here, "make_words_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(here, "words_golden.npz"))
    f1, f2 = mod.example_features(int(g["seed"]), int(g["nfeatures"]))
    return f1, g["w1"].astype(np.int32), f2, g["w2"].astype(np.int32)


def test_words_matcher_reproduces_the_reference_known_answer():
    """opensfm/test/test_matching.py:50-70 with its own assertions (lowes_ratio 0.8, bow_num_checks 20)."""
    f1, w1, f2, w2 = words_known_answer_case()
    matches = mo.match_using_words(f1, w1, f2, w2[:, 0], 0.8, 20)
    assert len(matches) == len(f1)
    assert all(i == j for i, j in matches)
