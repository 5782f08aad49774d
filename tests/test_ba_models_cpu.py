"""The product's per-observation math (opensfm_b200/csrc/ba_models.cuh, compiled for the
host by g++) against the oracle's dual numbers — catches derivation bugs without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import ba_lm as oracle
from opensfm_b200 import ba_problem as bp

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_harness", "ba_models_host.cpp")
LIB = os.path.join(HERE, "cpu_harness", "_build", "libba_models_host.so")
HDR = os.path.join(HERE, "..", "opensfm_b200", "csrc", "ba_models.cuh")


@pytest.fixture(scope="module")
def hd():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", LIB, SRC])
    L = ctypes.CDLL(LIB)
    L.hd_loss.restype = ctypes.c_double
    return L


def _eval(L, t, cam, ri, rc, use, X, obs, sig):
    dp = ctypes.POINTER(ctypes.c_double)
    C = L.hd_num_params(t)
    arr = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    cam, ri, rc, X, obs = map(arr, (cam, ri, rc, X, obs))
    r, jc, ji, jr, jp = np.zeros(3), np.zeros(48), np.zeros(18), np.zeros(18), np.zeros(9)
    P = lambda x: x.ctypes.data_as(dp)
    n = L.hd_observation_eval(t, P(cam), P(ri), P(rc), int(use), P(X), P(obs), ctypes.c_double(sig), P(r), P(jc),
                              P(ji), P(jr), P(jp), 1)
    jrc = jr[:n * 6].reshape(n, 6) if use else np.zeros((n, 6))
    return r[:n], jc[:n * C].reshape(n, C), ji[:n * 6].reshape(n, 6), jrc, jp[:n * 3].reshape(n, 3)


def test_num_params(hd):
    assert [hd.hd_num_params(t) for t in range(10)] == [bp.camera_num_params(t) for t in range(10)]


def test_reference_vectors(hd):
    from test_oracle_ba import CAMS, OBS, POINT, RT, SIGMA

    for t, c in CAMS.items():
        for use in (True, False):
            a = _eval(hd, t, c, RT, RT, use, POINT, OBS, SIGMA)
            b = oracle.reprojection(t, c, RT, RT, use, POINT, OBS, SIGMA, autodiff=True)
            assert max(np.abs(x - y).max() for x, y in zip(a, b)) < 1e-14, (t, use)


def test_random_configurations(hd):
    rng = np.random.RandomState(0)
    focal_at = {0: 2, 1: 5, 2: 2, 3: 4, 4: 8, 5: 12, 7: 3, 8: 2, 9: 1}
    worst = 0.0
    for _ in range(1500):
        t = rng.randint(0, 10)
        C = hd.hd_num_params(t)
        cam = rng.normal(0, 0.02, C)
        if t != 6:
            cam[focal_at[t]] = 0.5 + rng.rand()
            if t in (1, 3, 4, 5, 8, 9):
                cam[focal_at[t] + 1] = 0.9 + 0.2 * rng.rand()
            if t == 7:
                cam[0] = rng.rand()
        sc = rng.choice([1.0, 1e-3, 1e-9, 0.0])
        ri = np.concatenate([rng.normal(0, 1, 3) * sc, rng.normal(0, 1, 3)])
        rc = np.concatenate([rng.normal(0, 0.3, 3) * sc, rng.normal(0, 0.2, 3)])
        X = rng.normal(0, 1, 3) + ri[3:] + np.array([0, 0, 4.0])
        use = bool(rng.randint(2))
        a = _eval(hd, t, cam, ri, rc, use, X, [0.1, -0.2], 0.004)
        b = oracle.reprojection(t, cam, ri, rc, use, X, [0.1, -0.2], 0.004, autodiff=True)
        scale = max(1.0, max(np.abs(y).max() for y in b))
        worst = max(worst, max(np.abs(x - y).max() for x, y in zip(a, b)) / scale)
    assert worst < 1e-11


def test_loss_matches_oracle(hd):
    w = ctypes.c_double()
    for i, name in enumerate(["TrivialLoss", "HuberLoss", "SoftLOneLoss", "CauchyLoss", "ArctanLoss"]):
        for s in (0.0, 0.3, 1.0, 7.0, 1e4):
            r0 = hd.hd_loss(i, ctypes.c_double(1.3), ctypes.c_double(s), ctypes.byref(w))
            ref = oracle.loss(name, 1.3, s)
            assert abs(r0 - ref[0]) < 1e-12 * max(1, abs(ref[0]))
            assert abs(w.value - np.sqrt(ref[1])) < 1e-14
