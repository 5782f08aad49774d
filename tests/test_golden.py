"""Committed golden fixtures (tests/golden/, made by make_golden.py from the reference matcher and
the BA oracle): the oracle on CPU, the CUDA path on the GPU."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

from oracle import ba_lm, match_oracle as mo  # noqa: E402

CFG = {"lowes_ratio": 0.8}
MATCH = np.load(os.path.join(HERE, "golden", "match_golden.npz"))
BA = np.load(os.path.join(HERE, "golden", "ba_golden.npz"))


def test_numpy_restatement_reproduces_cv2_fixtures():
    for name, f1, f2, mask in make_golden.match_cases():
        got = np.array(mo.match_brute_force_numpy(f1, f2, CFG, mask), dtype=np.int32).reshape(-1, 2)
        assert np.array_equal(got, MATCH[name + "_oneway"]), name
        sym = np.array(sorted(mo.match_brute_force_symmetric_numpy(f1, f2, CFG, mask)), dtype=np.int32).reshape(-1, 2)
        assert np.array_equal(sym, MATCH[name + "_sym"]), name


def test_ba_oracle_reproduces_fixture():
    res = ba_lm.solve(make_golden.ba_case())
    assert res["iterations"] == int(BA["iterations"])
    assert abs(res["final_cost"] - float(BA["final_cost"])) <= 1e-9 * float(BA["final_cost"])
    assert np.abs(res["points"] - BA["points"]).max() < 1e-9


@pytest.mark.gpu
def test_gpu_matcher_reproduces_cv2_fixtures():
    from opensfm_b200 import matching

    for name, f1, f2, mask in make_golden.match_cases():
        got = np.array(matching.match_brute_force(f1, f2, CFG, mask), dtype=np.int32).reshape(-1, 2)
        assert np.array_equal(got, MATCH[name + "_oneway"]), name
        sym = np.array(sorted(matching.match_brute_force_symmetric(f1, f2, CFG, mask)), dtype=np.int32).reshape(-1, 2)
        assert np.array_equal(sym, MATCH[name + "_sym"]), name


@pytest.mark.gpu
def test_gpu_ba_reproduces_fixture():
    from opensfm_b200 import bundle

    got = bundle.solve(make_golden.ba_case())
    assert abs(got["summary"]["final_cost"] - float(BA["final_cost"])) <= 1e-6 * float(BA["final_cost"])
    assert np.abs(got["points"] - BA["points"]).max() < 2e-5
    assert np.abs(got["inst"] - BA["inst"]).max() < 2e-5
    rm = lambda e: np.sqrt((e ** 2).sum(1).mean())
    assert abs(rm(got["reprojection_errors"]) - rm(BA["reprojection_errors"])) <= 1e-6 * rm(BA["reprojection_errors"])
