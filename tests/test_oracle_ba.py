"""Pins the BA oracle against the reference's own golden vectors / known-answer tests."""
import math

import numpy as np
import pytest

from oracle import ba_lm as oracle
from opensfm_b200 import ba_problem as bp
from opensfm_b200 import synthetic as syn

# opensfm/src/bundle/test/reprojection_errors_test.cc:10-18,105-174
POINT = [1.0, 2.0, 3.0]
RT = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6]
OBS = [0.5, 0.5]
SIGMA = 10.0
CAMS = {
    bp.BROWN: [0.3, 1.0, 0.001, -0.02, 0.1, -0.03, 0.001, -0.005, 0.001],
    bp.PERSPECTIVE: [0.3, 0.1, -0.03],
    bp.FISHEYE: [0.3, 0.1, -0.03],
    bp.FISHEYE_OPENCV: [0.3, 1.0, 0.001, -0.02, 0.1, -0.03, 0.001, -0.005],
    bp.FISHEYE62: [0.3, 1.0, 0.001, -0.02, 0.1, -0.03, 0.001, -0.005, 0.01, 0.006, 0.02, 0.003],
    bp.FISHEYE624: [0.3, 1.0, 0.001, -0.02, 0.1, -0.03, 0.001, -0.005, 0.01, 0.006, 0.02, 0.003, 0.001, -0.009,
                    -0.01, 0.03],
    bp.DUAL: [0.5, 0.3, 0.1, -0.03],
    bp.SPHERICAL: [0.0],
    bp.RADIAL: [0.1, -0.03, 0.3, 1.0, 0.001, -0.02],
    bp.SIMPLE_RADIAL: [0.1, 0.3, 1.0, 0.001, -0.02],
}


@pytest.mark.parametrize("ptype", sorted(CAMS))
@pytest.mark.parametrize("use_rc", [True, False])
def test_analytic_equals_autodiff_reference_tolerance(ptype, use_rc):
    """reprojection_errors_test.cc:52-80: analytic vs autodiff, eps = 1e-14."""
    a = oracle.reprojection(ptype, CAMS[ptype], RT, RT, use_rc, POINT, OBS, SIGMA, autodiff=False)
    b = oracle.reprojection(ptype, CAMS[ptype], RT, RT, use_rc, POINT, OBS, SIGMA, autodiff=True)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() < 1e-14


def test_fisheye624_golden_pixels():
    """geometry/test/camera_test.cc:119-172 (isApprox 1e-5)."""
    f = 200 / (2 * math.pi)
    pts = [[10.0, 0, 0], [0, 10.0, 0], [1e-6, 1e-6, 10.0]]
    thin = [0, 0, 0, 0, 0, 0, 0, 0, 0.01, -0.007, -0.03, 0.0053]
    ref = np.array([[149.42887062, 48.67088846], [99.4288738, 98.67088528], [100.0, 50.0]])
    got = np.array([oracle.project(bp.FISHEYE624, thin + [f, 1.0, 100, 50], p) for p in pts])
    assert np.linalg.norm(got - ref) <= 1e-5 * min(np.linalg.norm(got), np.linalg.norm(ref))
    tan = [0, 0, 0, 0, 0, 0, 0.03, 0.01, 0.01, -0.007, -0.03, 0.0053]
    ref2 = np.array([[151.7850648, 51.02708265], [100.2142719, 105.7394678], [100.0, 50.0]])
    got2 = np.array([oracle.project(bp.FISHEYE624, tan + [f, 1.0, 100, 50], p) for p in pts])
    assert np.linalg.norm(got2 - ref2) <= 1e-5 * min(np.linalg.norm(got2), np.linalg.norm(ref2))


def test_parameter_order_matches_camera_cc():
    """geometry/src/camera.cc:9-178: perspective stores [k1, k2, focal]; focal scales the output."""
    p = np.array([0.2, -0.1, 2.0])
    a = oracle.project(bp.PERSPECTIVE, [0.0, 0.0, 0.5], p)
    assert np.allclose(a, 0.5 * p[:2] / p[2])
    b = oracle.project(bp.BROWN, [0, 0, 0, 0, 0, 0.5, 2.0, 0.1, -0.2], p)
    assert np.allclose(b, [0.5 * 0.1 + 0.1, 0.5 * 2.0 * -0.05 - 0.2])
    assert [oracle.camera_num_params(t) for t in range(10)] == [3, 9, 3, 8, 12, 16, 1, 4, 6, 5]
    assert [bp.camera_num_params(t) for t in range(10)] == [3, 9, 3, 8, 12, 16, 1, 4, 6, 5]


def test_pose_matches_angle_axis_rotation():
    """geometry/test/camera_functions_test.cc:142-163: pose forward == R(-r)(X - t)."""
    from scipy.spatial.transform import Rotation

    r, t = np.array(RT[:3]), np.array(RT[3:])
    expected = Rotation.from_rotvec(-r).apply(np.array(POINT) - t)
    # a perspective camera with f=1 and no distortion exposes x/z, y/z of the camera point
    got = oracle.project(bp.PERSPECTIVE, [0, 0, 1.0], expected)
    res = oracle.reprojection(bp.PERSPECTIVE, [0, 0, 1.0], RT, None, False, POINT, [0, 0], 1.0)[0]
    assert np.abs(res - got).max() < 1e-15


def test_ceres_loss_values():
    """ceres::LossFunction definitions (SURVEY.md §8c box)."""
    for a in (0.5, 1.0, 2.0):
        b = a * a
        for s in (0.0, 0.1, 1.0, 10.0):
            assert np.allclose(oracle.loss("TrivialLoss", a, s), [s, 1.0])
            assert np.allclose(oracle.loss("SoftLOneLoss", a, s), [2 * b * (math.sqrt(1 + s / b) - 1), 1 / math.sqrt(1 + s / b)])
            assert np.allclose(oracle.loss("CauchyLoss", a, s), [b * math.log(1 + s / b), 1 / (1 + s / b)])
            assert np.allclose(oracle.loss("ArctanLoss", a, s), [a * math.atan2(s, a), 1 / (1 + s * s / b)])
            h = [s, 1.0] if s <= b else [2 * a * math.sqrt(s) - b, a / math.sqrt(s)]
            assert np.allclose(oracle.loss("HuberLoss", a, s), h)


def test_lm_recovers_synthetic_scene():
    """opensfm/test/test_bundle.py:116-165: std of reprojection errors < 5e-3 after BA."""
    sc = syn.cube_scene(8, 500, 1.0, with_descriptors=False)
    pb = syn.scene_to_problem(sc)
    res = oracle.solve(pb)
    assert res["termination"] == "CONVERGENCE"
    assert res["final_cost"] < 0.05 * res["initial_cost"]
    err = res["reprojection_errors"][:, :2]
    assert err.std() < 5e-3
    # noise floor: 1 px / 800
    assert abs(np.sqrt((err ** 2).sum(1).mean()) - math.sqrt(2) / 800) < 3e-4


def test_lm_exact_data_reaches_ground_truth():
    sc = syn.cube_scene(6, 300, 0.0, with_descriptors=False)
    pb = syn.scene_to_problem(sc, loss_name="TrivialLoss")
    res = oracle.solve(pb)
    assert np.abs(res["reprojection_errors"]).max() < 1e-7


def test_unknown_loss_name_raises():
    sc = syn.cube_scene(3, 50, 1.0, with_descriptors=False)
    pb = syn.scene_to_problem(sc)
    pb.loss_name = "NopeLoss"
    with pytest.raises(RuntimeError):
        oracle.solve(pb)
