"""GPU parity of the brute-force matcher against the reference's cv2 path (bit-exact indices)."""
import numpy as np
import pytest

from oracle import match_oracle as mo
from opensfm_b200 import matching, synthetic as syn

pytestmark = pytest.mark.gpu
CFG = {"lowes_ratio": 0.8}


def _pairset(x):
    return sorted((int(a), int(b)) for a, b in x)


def _related(n1, n2, seed, lo=0, hi=255):
    a = syn.hahog_like_descriptors(n1, seed)
    b = syn.hahog_like_descriptors(n2, seed + 1)
    k = min(n1, n2) // 2
    rng = np.random.RandomState(seed + 2)
    b[:k] = np.clip(a[:k] + rng.randint(-6, 7, (k, 128)), lo, hi).astype(np.float32)
    perm = rng.permutation(n2)
    return a, b[perm]


@pytest.mark.parametrize("n1,n2", [(1, 2), (2, 1), (5, 3), (64, 64), (65, 129), (300, 1000), (1500, 1700)])
def test_l2_one_way_matches_cv2(n1, n2):
    a, b = _related(n1, n2, 10 + n1)
    assert matching.match_brute_force(a, b, CFG) == mo.match_brute_force(a, b, CFG)


def test_l2_ties_lowest_index_first():
    a, b = _related(200, 400, 3)
    b[300:310] = b[0:10]  # exact duplicates -> ties
    for ratio in (0.8, 1.0, 1.5):
        cfg = {"lowes_ratio": ratio}
        assert matching.match_brute_force(a, b, cfg) == mo.match_brute_force(a, b, cfg)


def test_l2_general_float_descriptors_simt():
    rng = np.random.RandomState(0)
    a = rng.rand(500, 128).astype(np.float32)
    b = rng.rand(700, 128).astype(np.float32)
    b[:250] = a[:250] + rng.normal(0, 0.02, (250, 128)).astype(np.float32)
    assert matching.match_brute_force(a, b, CFG) == mo.match_brute_force(a, b, CFG)


@pytest.mark.parametrize("n1,n2,dim", [(8000, 8000, 128), (3000, 4000, 64), (700, 900, 100), (500, 600, 20),
                                       (400, 500, 384), (300, 400, 700)])
def test_l2_general_float_exact_at_scale(n1, n2, dim):
    """Arbitrary float descriptors (e.g. root-SIFT): bit-exact match lists vs live cv2, including near-ties --
    the kernel sums in cv2's order (match.cu bf_top2_f32_cv).  8000 x 8000 is the C3 descriptor count."""
    rng = np.random.RandomState(n1 + dim)
    a = rng.rand(n1, dim).astype(np.float32)
    b = rng.rand(n2, dim).astype(np.float32)
    k = min(n1, n2) // 2
    b[:k] = a[:k] + rng.normal(0, 0.02, (k, dim)).astype(np.float32)
    b[k:k + 50] = b[:50] + np.float32(1e-7)      # near-duplicate trains: order decided by the last bits
    for ratio in (0.8, 1.0):
        cfg = {"lowes_ratio": ratio}
        assert matching.match_brute_force(a, b, cfg) == mo.match_brute_force(a, b, cfg)
    assert _pairset(matching.match_brute_force_symmetric(a, b, CFG)) == _pairset(
        mo.match_brute_force_symmetric(a, b, CFG))


def test_l2_integer_descriptors_16000_rows():
    a, b = _related(16000, 15000, 77)
    assert matching.match_brute_force(a, b, CFG) == mo.match_brute_force(a, b, CFG)


def test_l2_float_too_long_raises():
    a = np.zeros((4, 800), np.float32) + 0.5
    with pytest.raises((RuntimeError, ValueError)):
        matching.match_brute_force(a, a, CFG)


@pytest.mark.parametrize("dim", [32, 64, 100, 128])
def test_l2_other_dims(dim):
    rng = np.random.RandomState(dim)
    a = rng.randint(0, 256, (333, dim)).astype(np.float32)
    b = rng.randint(0, 256, (444, dim)).astype(np.float32)
    b[:100] = np.clip(a[:100] + rng.randint(-3, 4, (100, dim)), 0, 255)
    assert matching.match_brute_force(a, b, CFG) == mo.match_brute_force(a, b, CFG)


def test_l2_mask():
    a, b = _related(400, 500, 7)
    rng = np.random.RandomState(1)
    mask = rng.rand(400, 500) < 0.05
    mask[:10] = False  # queries with no candidate
    mask[10, :] = False
    mask[10, 3] = True  # exactly one candidate -> dropped (matching.py:752)
    assert matching.match_brute_force(a, b, CFG, mask) == mo.match_brute_force(a, b, CFG, mask)
    assert _pairset(matching.match_brute_force_symmetric(a, b, CFG, mask)) == _pairset(
        mo.match_brute_force_symmetric(a, b, CFG, mask))


def test_l2_symmetric():
    a, b = _related(900, 1100, 21)
    assert _pairset(matching.match_brute_force_symmetric(a, b, CFG)) == _pairset(
        mo.match_brute_force_symmetric(a, b, CFG))


@pytest.mark.parametrize("nbytes", [32, 61, 64])
def test_hamming(nbytes):
    u1 = syn.binary_descriptors(700, 1, nbytes)
    u2 = syn.binary_descriptors(900, 2, nbytes)
    u2[:300] = u1[:300]
    u2[:300, :3] ^= 5
    assert matching.match_brute_force(u1, u2, CFG) == mo.match_brute_force(u1, u2, CFG)
    assert _pairset(matching.match_brute_force_symmetric(u1, u2, CFG)) == _pairset(
        mo.match_brute_force_symmetric(u1, u2, CFG))


@pytest.mark.parametrize("nbytes,n1,n2", [(61, 3000, 2700), (32, 1234, 4321), (4, 600, 700), (63, 130, 129)])
def test_hamming_tensor_core_kernel(nbytes, n1, n2):
    """Binary descriptors with at most 63 bytes run as a +-1 fp8 contraction on the tensor cores (kernel id 3);
    results are cv2's, ties included (4-byte descriptors: nearly every row has tied candidates), and equal to the
    popcount kernel's."""
    u1 = syn.binary_descriptors(n1, 11, nbytes)
    u2 = syn.binary_descriptors(n2, 12, nbytes)
    k = min(n1, n2) // 3
    u2[:k] = u1[:k]
    u2[:k, :2] ^= 9
    got = {}
    for kernel in (0, 1):
        pm = matching.PairMatcher(kernel=kernel)
        pm.add("a", u1)
        pm.add("b", u2)
        got[kernel] = pm.match_pairs([("a", "b")], CFG, symmetric=False)[("a", "b")]
        assert pm.last_kernel() == (3 if kernel == 0 else 1)
        sym = pm.match_pairs([("a", "b")], CFG, symmetric=True)[("a", "b")]
        assert _pairset(sym) == _pairset(mo.match_brute_force_symmetric(u1, u2, CFG))
    want = np.asarray(mo.match_brute_force(u1, u2, CFG), dtype=np.int64).reshape(-1, 2)
    assert np.array_equal(np.asarray(got[0], dtype=np.int64).reshape(-1, 2), want)
    assert np.array_equal(np.asarray(got[1], dtype=np.int64).reshape(-1, 2), want)


def test_replacing_descriptor_sets_key_by_key_reuses_device_memory():
    """A long-lived matcher that replaces the descriptors of its keys one at a time (add allocates the new set before
    the old one is removed) must not grow: the slab allocator reuses released ranges."""
    pm = matching.PairMatcher()
    descs = [syn.hahog_like_descriptors(3000, 100 + i) for i in range(6)]
    for i, d in enumerate(descs):
        pm.add(i, d)
    reserved0, used0 = pm.device_bytes()
    for rnd in range(40):
        k = rnd % 6
        pm.add(k, syn.hahog_like_descriptors(3000, 1000 + rnd))
    reserved1, used1 = pm.device_bytes()
    assert reserved1 == reserved0, (reserved0, reserved1)
    assert used1 <= used0 + used0 // 4
    a, b = syn.hahog_like_descriptors(3000, 1000 + 36), syn.hahog_like_descriptors(3000, 1000 + 37)
    assert _pairset(pm.match_pairs([(0, 1)], CFG)[(0, 1)]) == _pairset(mo.match_brute_force_symmetric(a, b, CFG))


def test_empty_inputs():
    a = np.zeros((0, 128), np.float32)
    b = syn.hahog_like_descriptors(10, 1)
    assert matching.match_brute_force(a, b, CFG) == []
    assert matching.match_brute_force(b, a, CFG) == []


def test_dtype_mismatch_asserts():
    with pytest.raises(AssertionError):
        matching.match_brute_force(np.zeros((2, 8), np.float32), np.zeros((2, 8), np.uint8), CFG)


@pytest.mark.parametrize("kernel", [1, 2])
def test_pair_batch_cube_scene_both_kernels(kernel):
    sc = syn.cube_scene(6, 1500, 1.0)
    feats = {s: sc.features_of_shot(s)[0] for s in range(6)}
    pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)]
    pm = matching.PairMatcher(kernel=kernel)
    for s, f in feats.items():
        pm.add(s, f)
    res = pm.match_pairs(pairs, {"lowes_ratio": 0.8, "symmetric_matching": True})
    assert pm.last_kernel() == kernel
    for (i, j) in pairs:
        ref = _pairset(mo.match_brute_force_symmetric(feats[i], feats[j], CFG))
        assert _pairset(res[(i, j)]) == ref, (i, j)
    res1 = pm.match_pairs(pairs, {"lowes_ratio": 0.8}, symmetric=False)
    for (i, j) in pairs:
        assert [tuple(x) for x in res1[(i, j)].tolist()] == mo.match_brute_force(feats[i], feats[j], CFG)


def test_large_norms_route_to_exact_kernel():
    """d^2 near 2^23: distinct d^2 collapse to one float32 sqrt; cv2 ranks on the sqrt."""
    rng = np.random.RandomState(5)
    a = rng.randint(200, 256, (600, 128)).astype(np.float32)
    b = rng.randint(0, 40, (700, 128)).astype(np.float32)
    # norms this large break the d^2 < 2^22 guarantee: automatic selection must fall to the exact
    # SIMT kernel, and forcing the tensor-core kernel must be refused
    pm = matching.PairMatcher(kernel=0)
    pm.add("a", a)
    pm.add("b", b)
    for ratio in (0.999, 1.0, 1.01):
        got = pm.match_pairs([("a", "b")], {"lowes_ratio": ratio}, symmetric=False)[("a", "b")]
        assert pm.last_kernel() == 1
        assert [tuple(x) for x in got.tolist()] == mo.match_brute_force(a, b, {"lowes_ratio": ratio})
    pm2 = matching.PairMatcher(kernel=2)
    pm2.add("a", a)
    pm2.add("b", b)
    with pytest.raises(ValueError):
        pm2.match_pairs([("a", "b")], {"lowes_ratio": 0.8}, symmetric=False)


def test_threads_share_nothing():
    import threading

    a, b = _related(500, 600, 33)
    ref = mo.match_brute_force(a, b, CFG)
    out = [None] * 4

    def work(i):
        out[i] = matching.match_brute_force(a, b, CFG)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(o == ref for o in out)


def test_uint8_stored_l2_descriptors_match_their_float32_form():
    """HAHOG / SIFT-uchar uploaded as bytes (osfm_matcher_add_u8_l2) give exactly the float32 results."""
    a, b = _related(1500, 1700, 31)
    pm = matching.PairMatcher()
    pm.add("a", a.astype(np.uint8), uint8_is_l2=True)
    pm.add("b", b.astype(np.uint8), uint8_is_l2=True)
    got = pm.match_pairs([("a", "b")], CFG, symmetric=False)[("a", "b")]
    assert pm.last_kernel() == 2
    assert [tuple(x) for x in got.tolist()] == mo.match_brute_force(a, b, CFG)
    pm.add_many([("c", a.astype(np.uint8)), ("d", b.astype(np.uint8))], uint8_is_l2=True)
    sym = pm.match_pairs([("c", "d")], CFG)[("c", "d")]
    assert _pairset(sym) == _pairset(mo.match_brute_force_symmetric(a, b, CFG))
    # a short descriptor (no tensor-core operands) goes through the widening kernel
    a20, b20 = a[:300, :20], b[:400, :20]
    pm.add("e", a20.astype(np.uint8), uint8_is_l2=True)
    pm.add("f", b20.astype(np.uint8), uint8_is_l2=True)
    got20 = pm.match_pairs([("e", "f")], CFG, symmetric=False)[("e", "f")]
    assert [tuple(x) for x in got20.tolist()] == mo.match_brute_force(np.ascontiguousarray(a20), np.ascontiguousarray(b20), CFG)


@pytest.mark.parametrize("n_desc,kernel", [(900, 0), (900, 1), (2500, 0)])
def test_guided_matching_matches_the_reference_mask_and_matcher(n_desc, kernel):
    """matching._match_descriptors_guided_impl (matching.py:260-338): epipolar mask from bearings + relative pose
    (fp64, built on the device as a bitmask) and symmetric brute-force matching under it -- against the numpy
    restatement of EpipolarAngleTwoBearingsMany + live cv2 with the byte mask."""
    descs, bears, Rs, Os = syn.guided_scene(4, n_desc, seed=n_desc)
    pm = matching.PairMatcher(kernel=kernel)
    for i in range(4):
        pm.add(i, descs[i])
        pm.set_bearings(i, bears[i])
    pairs = [(0, 1), (1, 2), (0, 3), (3, 2)]
    poses = [syn.relative_pose(Rs[a], Os[a], Rs[b], Os[b]) for a, b in pairs]
    thr = 0.006   # config.py guided_matching_threshold
    got = pm.match_pairs_guided(pairs, poses, thr, CFG, mask_budget_bytes=3 * n_desc * n_desc // 8 * 2)
    assert pm.last_kernel() == (1 if kernel == 1 else 2)
    total = 0
    for (a, b), (R, t) in zip(pairs, poses):
        mask = mo.epipolar_mask(bears[a], bears[b], R, t, thr)
        ref = mo.match_brute_force_symmetric(descs[a], descs[b], CFG, mask)
        assert _pairset(got[(a, b)]) == _pairset(ref), (a, b)
        total += len(ref)
        assert 0.001 < mask.mean() < 0.2
    assert total > 50


def test_match_images_with_pairs_driver_applies_the_reference_gates():
    """matching.match (matching.py:563-634): min-match gate, robust filter hand-off, unfilter_matches."""
    rng = np.random.RandomState(3)
    feats, masks = {}, {}
    base = syn.hahog_like_descriptors(900, 400)
    for i in range(4):
        idx = rng.choice(900, 600, replace=False)
        feats["im%d" % i] = np.clip(base[idx] + rng.randint(-3, 4, (600, 128)), 0, 255).astype(np.float32)
        m = np.zeros(800, dtype=bool)
        m[rng.choice(800, 600, replace=False)] = True
        masks["im%d" % i] = m
    feats["junk"] = syn.hahog_like_descriptors(600, 999)          # shares nothing: fails the min-match gate
    masks["junk"] = np.ones(600, dtype=bool)
    pairs = [("im0", "im1"), ("im1", "im2"), ("im0", "junk"), ("im2", "im3")]
    cfg = dict(CFG, robust_matching_min_match=20, symmetric_matching=True)
    calls = []

    def robust(im1, im2, m):
        calls.append((im1, im2))
        return m[::2]                                              # stand-in for the geometric verification

    got = matching.match_images_with_pairs(feats, pairs, cfg, robust_filter=robust, feature_masks=masks)
    for p in pairs:
        ref = np.array(sorted(mo.match_brute_force_symmetric(feats[p[0]], feats[p[1]], cfg)), dtype=np.int64).reshape(-1, 2)
        if len(ref) < 20:
            assert p == ("im0", "junk") and len(got[p]) == 0 and p not in calls
            continue
        assert p in calls
        want = matching.unfilter_matches(np.array(sorted(map(tuple, got_raw(p, feats, cfg))))[::2], masks[p[0]], masks[p[1]])
        assert np.array_equal(np.array(sorted(map(tuple, got[p].tolist()))), np.array(sorted(map(tuple, want.tolist()))))


def got_raw(p, feats, cfg):
    # the device's own symmetric list in its (query-ordered) order: what the robust filter stand-in received
    pm = matching.PairMatcher()
    pm.add_many([(k, feats[k]) for k in p])
    return pm.match_pairs([p], cfg)[p].tolist()


@pytest.mark.parametrize("integer", [True, False])
def test_words_matcher_matches_the_restated_reference(integer):
    """features::match_using_words (features/src/matching.cc:24-88) incl. its candidate order, the max_checks break
    and the single-candidate case."""
    rng = np.random.RandomState(5)
    n1, n2, vocab, k = 700, 800, 120, 5
    if integer:
        a, b = _related(n1, n2, 55)
    else:
        a = rng.rand(n1, 128).astype(np.float32)
        b = rng.rand(n2, 128).astype(np.float32)
        b[:300] = a[:300] + rng.normal(0, 0.02, (300, 128)).astype(np.float32)
    centers = a[rng.choice(n1, vocab, replace=False)]

    def nearest_words(f, kk):
        d = ((f[:, None, :].astype(np.float64) - centers[None].astype(np.float64)) ** 2).sum(2)
        return np.argsort(d, axis=1, kind="stable")[:, :kk].astype(np.int32)

    w1, w2 = nearest_words(a, k), nearest_words(b, k)
    for checks in (20, 3):
        cfg = {"lowes_ratio": 0.8, "bow_num_checks": checks}
        got = matching.match_words(a, w1, b, w2, cfg)
        ref = mo.match_using_words(a, w1, b, w2[:, 0], 0.8, checks)
        assert np.array_equal(got, ref)
    sym = matching.match_words_symmetric(a, w1, b, w2, {"lowes_ratio": 0.8, "bow_num_checks": 20})
    r12 = {tuple(x) for x in mo.match_using_words(a, w1, b, w2[:, 0], 0.8, 20).tolist()}
    r21 = {(y, x) for x, y in mo.match_using_words(b, w2, a, w1[:, 0], 0.8, 20).tolist()}
    assert set(sym) == (r12 & r21) and len(sym) > 20


def test_words_matcher_reproduces_the_reference_known_answer():
    """opensfm/test/test_matching.py:50-70 on the CUDA kernel: every feature matches its noisy copy."""
    from test_match_oracle import words_known_answer_case

    f1, w1, f2, w2 = words_known_answer_case()
    matches = matching.match_words(f1, w1, f2, w2, {"lowes_ratio": 0.8, "bow_num_checks": 20})
    assert len(matches) == len(f1)
    assert all(i == j for i, j in matches)
    assert np.array_equal(matches, mo.match_using_words(f1, w1, f2, w2[:, 0], 0.8, 20))


def test_vlad_distances():
    rng = np.random.RandomState(1)
    hist = {"im%d" % i: rng.normal(0, 1, 64 * 128).astype(np.float32) for i in range(9)}
    im, dist, others = matching.vlad_distances("im3", ["im%d" % i for i in range(9)] + ["missing"], hist)
    assert im == "im3" and "im3" not in others and "missing" not in others and len(others) == 8
    want = [float(np.linalg.norm(hist["im3"].astype(np.float64) - hist[o].astype(np.float64))) for o in others]
    assert np.allclose(dist, want, rtol=1e-6)
    assert matching.vlad_distances("nope", ["im0"], hist) == ("nope", [], [])
