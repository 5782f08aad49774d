"""B-py-1: `opensfm_b200.reconstruction.{bundle, bundle_shot_poses, bundle_local}` as drop-ins of
`opensfm.reconstruction.*` (reconstruction.py:69-126 = sfm::BAHelpers, ba_helpers.cc:117-819).

Each test runs on the oracle (CPU: `bundle.solve` swapped for the oracle's LM, which exercises the whole host logic
-- neighbourhoods, fixed / free masks, priors, write-back, report) and on the CUDA engine (gpu marker).  The three
reference tests that go through `reconstruction.bundle()` are ported with their own tolerances:
  test_bundle_projection_fixed_internals  opensfm/test/test_bundle.py:116-165
  test_bundle_void_gps_ignored            :641-685
  test_bundle_alignment_prior             :688-716"""
import copy

import numpy as np
import pytest

from oracle import ba_lm as oracle
from opensfm_b200 import bundle as obundle
from opensfm_b200 import map_types as M
from opensfm_b200 import reconstruction as orec
from opensfm_b200 import synthetic as syn
from opensfm_b200 import types as T

BACKENDS = [pytest.param("oracle", id="oracle"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]

# opensfm/config.py defaults of every key the three functions read
CONFIG = {
    "bundle_analytic_derivatives": True, "loss_function": "SoftLOneLoss", "loss_function_threshold": 1,
    "exif_focal_sd": 0.01, "aspect_ratio_sd": 0.01, "principal_point_sd": 0.01, "radial_distortion_k1_sd": 0.01,
    "radial_distortion_k2_sd": 0.01, "radial_distortion_k3_sd": 0.01, "radial_distortion_k4_sd": 0.01,
    "tangential_distortion_p1_sd": 0.01, "tangential_distortion_p2_sd": 0.01, "gcp_horizontal_sd": 0.01,
    "gcp_vertical_sd": 0.1, "gcp_global_weight": 0.01, "rig_translation_sd": 0.1, "rig_rotation_sd": 0.1,
    "bundle_outlier_filtering_type": "FIXED", "bundle_outlier_auto_ratio": 3.0, "bundle_outlier_fixed_threshold": 0.006,
    "optimize_camera_parameters": True, "bundle_max_iterations": 100, "local_bundle_radius": 3,
    "local_bundle_min_common_points": 20, "local_bundle_max_shots": 30, "align_method": "auto",
    "align_orientation_prior": "horizontal", "bundle_use_gps": True, "bundle_use_gcp": True,
    "bundle_compensate_gps_bias": False, "processes": 1,
}


def _oracle_solve(pb, device=0, compute_reprojection_errors=True, **kw):
    res = oracle.solve(pb)
    res["summary"] = {"iterations": res["iterations"], "initial_cost": res["initial_cost"],
                      "final_cost": res["final_cost"], "termination": res["termination"]}
    return res


@pytest.fixture()
def backend(request, monkeypatch):
    if request.param == "oracle":
        monkeypatch.setattr(obundle, "solve", _oracle_solve)
    return request.param


def _scene_reconstruction(num_cameras=12, num_points=400, noise=1.0, max_obs=None, seed=42, perturb=True):
    """types.Reconstruction-like map of the synthetic cube scene, one perspective camera shared by all shots."""
    sc = syn.cube_scene(num_cameras, num_points, noise, seed=seed, with_descriptors=False, max_obs_per_point=max_obs)
    pb = syn.scene_to_problem(sc, shared_intrinsics=True) if perturb else syn.scene_to_problem(sc, perturb_seed=None, shared_intrinsics=True)
    r = M.Reconstruction()
    cam = T.Camera.create_perspective(float(sc.cam_params[0][2]), float(sc.cam_params[0][0]), float(sc.cam_params[0][1]))
    cam.id = "1"
    r.add_camera(cam)
    for s in range(sc.num_shots):
        r.create_shot("shot%d" % s, "1", T.Pose.from_ba_params(pb.inst[s]))
    for p in range(len(pb.points)):
        r.create_point(str(p), pb.points[p])
    for k in range(pb.num_observations):
        r.add_observation("shot%d" % pb.obs_shot[k], str(pb.obs_point[k]),
                          M.Observation(pb.obs_xy[k, 0], pb.obs_xy[k, 1], pb.obs_sigma[k], feature=k))
    return r, sc


def _errors_std(points):
    all_errors = []
    for p in points.values():
        all_errors += list(p.reprojection_errors.values())
    return float(np.std(all_errors))


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
def test_bundle_projection_fixed_internals(backend):
    reference, _ = _scene_reconstruction(10, 500, 1.0)
    camera_priors = dict(reference.cameras.items())
    rig_priors = dict(reference.rig_cameras.items())
    orig = copy.deepcopy(reference.cameras["1"])
    cfg = dict(CONFIG, bundle_use_gps=False, optimize_camera_parameters=False)
    report = orec.bundle(reference, camera_priors, rig_priors, [], cfg)
    assert _errors_std(reference.points) < 5e-3
    assert reference.cameras["1"].focal == orig.focal
    assert reference.cameras["1"].k1 == orig.k1 and reference.cameras["1"].k2 == orig.k2
    # report keys consumed by reconstruction.log_bundle_stats (reconstruction.py:52-66)
    assert set(report) >= {"brief_report", "wall_times", "num_images", "num_points", "num_reprojections"}
    assert set(report["wall_times"]) == {"setup", "run", "teardown"}
    assert report["num_images"] == 10 and report["num_points"] == len(reference.points)
    lm = next(iter(reference.points.values()))
    assert all(len(e) == 2 for e in lm.reprojection_errors.values()) and len(lm.reprojection_errors) > 0


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
def test_bundle_void_gps_ignored(backend):
    camera = T.Camera.create_perspective(1.0, 0.0, 0.0)
    camera.id = "camera1"
    r = M.Reconstruction()
    r.add_camera(camera)
    rng = np.random.RandomState(3)
    shot = r.create_shot("1", camera.id, T.Pose(rng.rand(3), rng.rand(3)))
    camera_priors = {camera.id: camera}
    rig_priors = dict(r.rig_cameras.items())
    # missing position
    shot.metadata.gps_position.value = np.zeros(3)
    shot.metadata.gps_accuracy.value = 1
    shot.metadata.gps_position.reset()
    shot.pose.set_origin(np.ones(3))
    orec.bundle(r, camera_priors, rig_priors, [], CONFIG)
    assert np.allclose(shot.pose.get_origin(), np.ones(3))
    # missing accuracy
    shot.metadata.gps_position.value = np.zeros(3)
    shot.metadata.gps_accuracy.value = 1
    shot.metadata.gps_accuracy.reset()
    shot.pose.set_origin(np.ones(3))
    orec.bundle(r, camera_priors, rig_priors, [], CONFIG)
    assert np.allclose(shot.pose.get_origin(), np.ones(3))
    # valid position and accuracy
    shot.metadata.gps_position.value = np.zeros(3)
    shot.metadata.gps_accuracy.value = 1
    shot.pose.set_origin(np.ones(3))
    orec.bundle(r, camera_priors, rig_priors, [], CONFIG)
    assert np.allclose(shot.pose.get_origin(), np.zeros(3))


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
def test_bundle_alignment_prior(backend):
    camera = T.Camera.create_perspective(1.0, 0.0, 0.0)
    camera.id = "camera1"
    r = M.Reconstruction()
    r.add_camera(camera)
    rng = np.random.RandomState(4)
    shot = r.create_shot("1", camera.id, T.Pose(rng.rand(3), rng.rand(3)))
    shot.metadata.gps_position.value = np.array([0, 0, 0])
    shot.metadata.gps_accuracy.value = 1
    orec.bundle(r, {camera.id: camera}, dict(r.rig_cameras.items()), [], CONFIG)
    shot = r.shots[shot.id]
    assert np.allclose(shot.pose.translation, np.zeros(3))
    # up vector in camera coordinates is (0, -1, 0)
    assert np.allclose(shot.pose.transform([0, 0, 1]), [0, -1, 0], atol=1e-7)


def test_shot_neighborhood_matches_the_reference_definition():
    r, _ = _scene_reconstruction(14, 600, 1.0, max_obs=4)
    interior, boundary = orec.shot_neighborhood_ids(r, "shot0", 2, 5, 6)
    assert "shot0" in interior and not (interior & boundary) and len(interior) <= 6 + 1
    # every boundary shot shares a point with the interior; nothing outside interior + boundary does
    ipts = {lm.id for s in interior for lm in r.shots[s].get_valid_landmarks()}
    for sid, shot in r.shots.items():
        shares = any(lm.id in ipts for lm in shot.get_valid_landmarks())
        if sid in boundary:
            assert shares
        elif sid not in interior:
            assert not shares
    # radius 1: the central shot alone (its rig instance)
    i1, _ = orec.shot_neighborhood_ids(r, "shot3", 1, 5, 30)
    assert i1 == {"shot3"}


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
def test_bundle_local_frees_interior_and_fixes_boundary(backend):
    r, _ = _scene_reconstruction(14, 600, 1.0, max_obs=4)
    before = {iid: T.pose_to_ba_params(inst.pose).copy() for iid, inst in r.rig_instances.items()}
    pts_before = {p: lm.coordinates.copy() for p, lm in r.points.items()}
    cfg = dict(CONFIG, bundle_use_gps=False, local_bundle_radius=2, local_bundle_min_common_points=5, local_bundle_max_shots=6)
    interior, boundary = orec.shot_neighborhood_ids(r, "shot0", 2, 5, 6)
    pt_ids, report = orec.bundle_local(r, dict(r.cameras.items()), dict(r.rig_cameras.items()), None, "shot0", cfg)
    moved = {iid for iid, inst in r.rig_instances.items() if not np.allclose(T.pose_to_ba_params(inst.pose), before[iid], atol=1e-12)}
    assert moved <= interior and len(moved) >= 1
    assert report["num_interior_images"] == len(interior) and report["num_boundary_images"] == len(boundary)
    assert report["num_other_images"] == len(r.shots) - len(interior) - len(boundary)
    assert report["num_points"] == len(pt_ids) > 0
    ipts = {lm.id for s in interior for lm in r.shots[s].get_valid_landmarks()}
    assert set(pt_ids) == ipts
    untouched = [p for p in r.points if p not in ipts]
    assert all(np.array_equal(r.points[p].coordinates, pts_before[p]) for p in untouched)
    assert all(r.points[p].reprojection_errors for p in pt_ids)


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
def test_bundle_shot_poses_moves_only_the_given_shots(backend):
    r, _ = _scene_reconstruction(8, 400, 1.0, perturb=False)
    rng = np.random.RandomState(0)
    truth = T.pose_to_ba_params(r.rig_instances["shot2"].pose).copy()
    r.shots["shot2"].pose.set_origin(r.shots["shot2"].pose.get_origin() + rng.normal(0, 0.05, 3))
    before = {iid: T.pose_to_ba_params(inst.pose).copy() for iid, inst in r.rig_instances.items()}
    pts_before = {p: lm.coordinates.copy() for p, lm in r.points.items()}
    cfg = dict(CONFIG, bundle_use_gps=False)
    report = orec.bundle_shot_poses(r, {"shot2"}, dict(r.cameras.items()), dict(r.rig_cameras.items()), cfg)
    assert set(report) == {"brief_report", "wall_times"}
    for iid, inst in r.rig_instances.items():
        if iid != "shot2":
            assert np.array_equal(T.pose_to_ba_params(inst.pose), before[iid])
    assert all(np.array_equal(r.points[p].coordinates, pts_before[p]) for p in r.points)
    assert np.abs(T.pose_to_ba_params(r.rig_instances["shot2"].pose) - truth).max() < 2e-3   # resectioned back


@pytest.mark.parametrize("backend", BACKENDS, indirect=True)
def test_gcp_point_prior_anchors_the_reconstruction(backend):
    """AddGCPToBundle (ba_helpers.cc:349-406): a GCP with lla and image observations becomes a free point with a prior."""
    r, sc = _scene_reconstruction(8, 300, 0.5)

    class Ref:  # TopocentricConverter stand-in: lla are already topocentric metres here
        def to_topocentric(self, lat, lon, alt):
            return np.array([lat, lon, alt])

    r.reference = Ref()
    gcps = []
    for g, x in enumerate([np.array([0.1, 0.2, -0.1]), np.array([-0.3, 0.1, 0.2]), np.array([0.2, -0.3, 0.3])]):
        pt = M.GroundControlPoint()
        pt.id = "g%d" % g
        pt.lla = {"latitude": x[0], "longitude": x[1], "altitude": x[2]}
        pt.has_altitude = True
        for s in range(4):
            xc = sc.R_wc[s] @ (x - sc.origins[s])
            px = syn.project_perspective(xc[None, :], *sc.cam_params[s])[0]
            pt.add_observation(M.GroundControlPointObservation("shot%d" % s, px))
        gcps.append(pt)
    cfg = dict(CONFIG, bundle_use_gps=False, align_method="naive", gcp_global_weight=1.0)
    report = orec.bundle(r, dict(r.cameras.items()), dict(r.rig_cameras.items()), gcps, cfg)
    assert _errors_std(r.points) < 5e-3
    assert "gcp-g0" not in r.points   # GCP points live in the bundle problem only
    assert report["num_points"] == len(r.points)


def test_remove_outliers_thresholds_reprojection_errors():
    r, _ = _scene_reconstruction(6, 100, 1.0)
    for lm in r.points.values():
        lm.reprojection_errors = {s.id: np.array([1e-4, 1e-4]) for s in lm.get_observations()}
    victim = next(iter(r.points.values()))
    shots = list(victim.get_observations())
    victim.reprojection_errors[shots[0].id] = np.array([0.1, 0.0])
    n_obs = victim.number_of_observations()
    assert orec.remove_outliers(r, CONFIG) == 1
    assert victim.number_of_observations() == n_obs - 1
    cfg = dict(CONFIG, bundle_outlier_filtering_type="AUTO")
    assert orec.get_actual_threshold(cfg, r.points) > 0


@pytest.mark.gpu
def test_monkey_patched_functions_give_the_same_result_as_solve():
    """re-pointing opensfm.reconstruction.bundle at this module.  The map-level function and a
    direct bundle.solve() of the equivalent SoA problem agree."""
    r, sc = _scene_reconstruction(10, 500, 1.0)
    cfg = dict(CONFIG, bundle_use_gps=False, align_method="naive")
    pb = syn.scene_to_problem(sc, shared_intrinsics=True)
    direct = obundle.solve(pb)
    orec.bundle(r, dict(r.cameras.items()), dict(r.rig_cameras.items()), [], cfg)
    pts = np.array([r.points[str(p)].coordinates for p in range(len(pb.points))])
    # same engine, different observation order (per shot vs per scene): fp64 atomics and the PCG stop differ in
    # the last digits
    assert np.abs(pts - direct["points"]).max() < 1e-6
    inst = np.array([T.pose_to_ba_params(r.rig_instances["shot%d" % s].pose) for s in range(10)])
    assert np.abs(inst - direct["inst"]).max() < 1e-6
