"""The reference's known-answer tests of `pybundle.ReconstructionAlignment`
(opensfm/test/test_reconstruction_alignment.py, all eight), same inputs and tolerances, on the oracle (CPU) and on
the CUDA engine."""
import numpy as np
import pytest

from oracle import ba_lm as oracle
from opensfm_b200 import types as T
from opensfm_b200.alignment import RARelativeMotionConstraint, ReconstructionAlignment

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def run(ra, backend):
    if backend == "cuda":
        ra.run()
        return
    pb = ra.to_problem()
    res = oracle.solve(pb)
    res["summary"] = {"iterations": res["iterations"], "initial_cost": res["initial_cost"],
                      "final_cost": res["final_cost"], "termination": res["termination"]}
    ra.apply_results(pb, res)


def get_shot_origin(shot):
    return T.Pose(np.array([shot.rx, shot.ry, shot.rz]), np.array([shot.tx, shot.ty, shot.tz])).get_origin()


def get_reconstruction_origin(r):
    s = r.scale
    return T.Pose(np.array([r.rx, r.ry, r.rz]), np.array([r.tx / s, r.ty / s, r.tz / s])).get_origin()


@pytest.mark.parametrize("backend", BACKENDS)
def test_single_shot(backend):
    ra = ReconstructionAlignment()
    ra.add_shot("1", 0.5, 0, 0, 0, 0, 0, False)
    ra.add_absolute_position_constraint("1", 1, 0, 0, 1)
    run(ra, backend)
    assert np.allclose(get_shot_origin(ra.get_shot("1")), [1, 0, 0], atol=1e-6)


@pytest.mark.parametrize("backend", BACKENDS)
def test_singleton_reconstruction(backend):
    ra = ReconstructionAlignment()
    ra.add_shot("1", 0, 0, 0, 0, 0, 0, False)
    ra.add_reconstruction("a", 0, 0, 0, 0, 0, 0, 4, False)
    ra.add_relative_motion_constraint(RARelativeMotionConstraint("a", "1", 0, 0, 0, -1, 0, 0))
    ra.add_absolute_position_constraint("1", 1, 0, 0, 1)
    run(ra, backend)
    assert np.allclose(get_shot_origin(ra.get_shot("1")), [1, 0, 0], atol=1e-6)


@pytest.mark.parametrize("backend", BACKENDS)
def test_pair(backend):
    ra = ReconstructionAlignment()
    ra.add_shot("1", 0, 0, 0, 0, 0, 0, False)
    ra.add_shot("2", 0, 0, 0, 0, 0, 0, False)
    ra.add_reconstruction("a", 0, 0, 0, 0, 0, 0, 4, False)
    ra.add_relative_motion_constraint(RARelativeMotionConstraint("a", "1", 0, 0, 0, 0, 0, 0))
    ra.add_relative_motion_constraint(RARelativeMotionConstraint("a", "2", 0, 0, 0, -1, 0, 0))
    ra.add_absolute_position_constraint("1", 1, 0, 0, 1)
    ra.add_absolute_position_constraint("2", 3, 0, 0, 1)
    run(ra, backend)
    rec_a = ra.get_reconstruction("a")
    assert np.allclose(get_shot_origin(ra.get_shot("1")), [1, 0, 0], atol=1e-6)
    assert np.allclose(get_shot_origin(ra.get_shot("2")), [3, 0, 0], atol=1e-6)
    assert np.allclose(get_reconstruction_origin(rec_a), [1, 0, 0], atol=1e-6)
    assert np.allclose(rec_a.scale, 0.5)


@pytest.mark.parametrize("backend", BACKENDS)
def test_two_shots_one_fixed(backend):
    ra = ReconstructionAlignment()
    ra.add_shot("1", 0, 0, 0, -1, 0, 0, True)
    ra.add_shot("2", 0, 0, 0, 0, 0, 0, False)
    ra.add_reconstruction("a", 0, 0, 0, 0, 0, 0, 1, False)
    ra.add_relative_motion_constraint(RARelativeMotionConstraint("a", "1", 0, 0, 0, 0, 0, 0))
    ra.add_relative_motion_constraint(RARelativeMotionConstraint("a", "2", 0, 0, 0, -1, 0, 0))
    ra.add_absolute_position_constraint("1", 100, 0, 0, 1)   # ignored: shot 1 is fixed
    ra.add_absolute_position_constraint("2", 3, 0, 0, 1)
    run(ra, backend)
    rec_a = ra.get_reconstruction("a")
    assert np.allclose(get_shot_origin(ra.get_shot("1")), [1, 0, 0], atol=1e-6)
    assert np.allclose(get_shot_origin(ra.get_shot("2")), [3, 0, 0], atol=1e-6)
    assert np.allclose(get_reconstruction_origin(rec_a), [1, 0, 0], atol=1e-6)
    assert np.allclose(rec_a.scale, 0.5)


@pytest.mark.parametrize("backend", BACKENDS)
def test_two_reconstructions_soft_alignment(backend):
    ra = ReconstructionAlignment()
    for s in "1234":
        ra.add_shot(s, 0, 0, 0, 0, 0, 0, False)
    ra.add_reconstruction("a", 0, 0, 0, 0, 0, 0, 1, False)
    for s, tx in (("1", 0), ("2", -1), ("3", -2)):
        ra.add_relative_motion_constraint(RARelativeMotionConstraint("a", s, 0, 0, 0, tx, 0, 0))
    ra.add_reconstruction("b", 0, 0, 0, 0, 0, 0, 1, False)
    for s, tx in (("2", 0), ("3", -1), ("4", -2)):
        ra.add_relative_motion_constraint(RARelativeMotionConstraint("b", s, 0, 0, 0, tx, 0, 0))
    ra.add_absolute_position_constraint("1", 1, 0, 0, 1)
    ra.add_absolute_position_constraint("2", 2, 0, 0, 1)
    run(ra, backend)
    for k, s in enumerate("1234"):
        assert np.allclose(get_shot_origin(ra.get_shot(s)), [k + 1, 0, 0], atol=1e-6)
    rec_a, rec_b = ra.get_reconstruction("a"), ra.get_reconstruction("b")
    assert np.allclose(get_reconstruction_origin(rec_a), [1, 0, 0], atol=1e-6)
    assert np.allclose(get_reconstruction_origin(rec_b), [2, 0, 0], atol=1e-6)
    assert np.allclose(rec_a.scale, 1) and np.allclose(rec_b.scale, 1)


def _two_rigid(ra, a_constant):
    ra.add_shot("a_1", 0, 0, 0, -1, 0, 0, True)
    ra.add_shot("a_2", 0, 0, 0, -2, 0, 0, True)
    ra.add_shot("a_3", 0, 0, 0, 0, 0, 0, True)
    ra.add_shot("a_4", 0, 0, 0, 0, -1, 0, True)
    ra.add_shot("a_5", 0, 0, 0, -1, 0, 0, True)
    ra.add_reconstruction("a", 0, 0, 0, 0, 0, 0, 1, a_constant)
    ra.add_shot("b_3", 0, 0, 0, -1, -1, 0, True)
    ra.add_shot("b_4", 0, 0, 0, -1, -2, 0, True)
    ra.add_shot("b_5", 0, 0, 0, -2, -1, 0, True)
    ra.add_shot("b_6", 0, 0, 0, -4, 0, 0, True)
    ra.add_shot("b_7", 0, 0, 0, -5, 0, 0, True)
    ra.add_reconstruction("b", 0, 0, 0, 0, 0, 0, 1, False)


@pytest.mark.parametrize("backend", BACKENDS)
def test_two_reconstructions_rigid_alignment(backend):
    ra = ReconstructionAlignment()
    _two_rigid(ra, False)
    ra.add_relative_absolute_position_constraint("a", "a_3", 5, 5, 0, 1)
    ra.add_relative_absolute_position_constraint("a", "a_4", 5, 6, 0, 1)
    ra.add_relative_absolute_position_constraint("b", "b_3", 5, 5, 0, 1)
    ra.add_relative_absolute_position_constraint("b", "b_4", 5, 6, 0, 1)
    run(ra, backend)
    rec_a, rec_b = ra.get_reconstruction("a"), ra.get_reconstruction("b")
    assert np.allclose(get_reconstruction_origin(rec_a), [5, 5, 0], atol=1e-6)
    assert np.allclose(get_reconstruction_origin(rec_b), [4, 4, 0], atol=1e-6)
    assert np.allclose(rec_a.scale, 1) and np.allclose(rec_b.scale, 1)


@pytest.mark.parametrize("backend", BACKENDS)
def test_two_reconstructions_common_camera(backend):
    ra = ReconstructionAlignment()
    _two_rigid(ra, True)
    ra.add_common_camera_constraint("a", "a_3", "b", "b_3", 1, 1)
    ra.add_common_camera_constraint("a", "a_4", "b", "b_4", 1, 1)
    ra.add_common_camera_constraint("a", "a_5", "b", "b_5", 1, 1)
    run(ra, backend)
    rec_a, rec_b = ra.get_reconstruction("a"), ra.get_reconstruction("b")
    assert np.allclose(get_reconstruction_origin(rec_a), [0, 0, 0], atol=1e-6)
    assert np.allclose(get_reconstruction_origin(rec_b), [-1, -1, 0], atol=1e-6)
    assert np.allclose(rec_a.scale, 1) and np.allclose(rec_b.scale, 1)


@pytest.mark.parametrize("backend", BACKENDS)
def test_common_points(backend):
    ra = ReconstructionAlignment()
    ra.add_reconstruction("a", 0, 0, 0, 0, 0, 0, 1, True)
    ra.add_reconstruction("b", 0, 0, 0, 0, 0, 0, 1, False)
    ra.add_common_point_constraint("a", 0, 0, 0, "b", -1, 0, 0, 1.0)
    ra.add_common_point_constraint("a", 1, 0, 0, "b", 0, 0, 0, 1.0)
    run(ra, backend)
    assert np.allclose(get_reconstruction_origin(ra.get_reconstruction("b")), [1, 0, 0], atol=1e-6)
