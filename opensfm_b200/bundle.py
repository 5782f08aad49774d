"""Drop-in bundle adjustment: the `opensfm.pybundle.BundleAdjuster` surface and a bulk
SoA entry, both running on the CUDA engine (opensfm_b200/csrc/ba.cu) through the C ABI.

Reference interfaces mirrored here:
* class `pybundle.BundleAdjuster` (opensfm/src/bundle/python/pybind.cc:45-117,
  stub opensfm/src/bundle/pybundle.pyi:34-187): same method names, argument
  meaning, defaults (CauchyLoss(1), 500 iterations, SPARSE_SCHUR,
  bundle_adjuster.cc:24-44) and error behaviour (missing ids ->
  RuntimeError "... doesn't exist.", unknown loss / solver names -> RuntimeError).
* `solve(problem)`: the bulk path that replaces O(N) string-keyed Add* calls
  (SURVEY.md §7 "String-keyed API"), fed by opensfm_b200.ba_problem.BAProblem.

No CPU fallback: every `run()` goes to the GPU library.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from . import ba_problem as bp
from . import types as T

_TERMINATION = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}

# One engine handle per (thread, device): its HBM workspaces (Jacobian planes, reduced system, ...)
# are kept between solves, so repeated bundle() calls do not pay cudaMalloc/cudaFree every time.
_tls = threading.local()


class _Handle:
    def __init__(self, device: int):
        self.L = _lib.load()
        self.h = ctypes.c_void_p()
        _lib.check(self.L.osfm_ba_create(int(device), ctypes.byref(self.h)))

    def __del__(self):
        try:
            self.L.osfm_ba_destroy(self.h)
        except Exception:
            pass


def _handle(device: int) -> "_Handle":
    cache = getattr(_tls, "handles", None)
    if cache is None:
        cache = _tls.handles = {}
    if device not in cache:
        cache[device] = _Handle(device)
    return cache[device]


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def solve(pb: bp.BAProblem, device: int = 0, rank: int = 0, world: int = 1, allreduce=None,
          stream: Optional[int] = None, compute_reprojection_errors: bool = True,
          out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, Any]:
    """Run the GPU bundle adjustment on a BAProblem.  Returns updated parameter arrays,
    unscaled reprojection errors (bundle_adjuster.cc:1196-1208) and the run summary.

    Multi-GPU: pass rank/world and either allreduce="nccl" (the library opens its own NCCL
    communicator; torch.distributed must be initialised, it only carries the 128-byte id once) or a
    callable `allreduce(ptr:int, count:int, stream:int) -> None` that sums `count` float64 at device
    pointer `ptr` across ranks (see opensfm_b200.dist; used with gloo in the CPU tests).

    `out` may hold preallocated C-contiguous float64 arrays "points" (P, 3) and "reprojection_errors"
    (N, 3) to receive the results (page-locked buffers make the device->host copy a plain DMA)."""
    pb.validate()
    L = _lib.load()
    h = _handle(int(device)).h
    if True:
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        K, NI, NR = len(pb.cam_type), len(pb.inst), len(pb.rigcam)
        S, P, N = len(pb.shot_inst), len(pb.points), len(pb.obs_shot)
        keep = [i32(pb.cam_type), f64(pb.cam_params), i32(pb.cam_const), f64(pb.cam_prior), f64(pb.cam_prior_sigma),
                i32(pb.cam_prior_log)]
        _lib.check(L.osfm_ba_set_cameras(h, K, *[_p(a) for a in keep]))
        k2 = [f64(pb.inst), i32(pb.inst_const), i32(pb.inst_has_prior), f64(pb.inst_prior_pos), f64(pb.inst_prior_std)]
        _lib.check(L.osfm_ba_set_rig_instances(h, NI, *[_p(a) for a in k2]))
        k3 = [f64(pb.rigcam), i32(pb.rigcam_const)]
        _lib.check(L.osfm_ba_set_rig_cameras(h, NR, *[_p(a) for a in k3]))
        k4 = [i32(pb.shot_inst), i32(pb.shot_cam), i32(pb.shot_rc), i32(pb.shot_use_rc)]
        _lib.check(L.osfm_ba_set_shots(h, S, *[_p(a) for a in k4]))
        k5 = [f64(pb.points), i32(pb.point_const)]
        _lib.check(L.osfm_ba_set_points(h, P, *[_p(a) for a in k5]))
        k6 = [i32(pb.obs_shot), i32(pb.obs_point), f64(pb.obs_xy), f64(pb.obs_sigma)]
        _lib.check(L.osfm_ba_set_observations(h, N, *[_p(a) for a in k6]))
        if pb.loss_name not in _lib.LOSS_IDS:
            raise RuntimeError("ceres::LossFunction with name %s not found." % pb.loss_name)  # bundle_adjuster.cc:427
        _lib.check(L.osfm_ba_set_options(h, _lib.LOSS_IDS[pb.loss_name], float(pb.loss_threshold),
                                         int(pb.max_iterations), pb.linear_solver_type.encode(),
                                         int(compute_reprojection_errors)))
        cb = None
        if world > 1:
            if allreduce is None:
                raise ValueError("world > 1 needs allreduce: a callable or the string 'nccl'")
            if isinstance(allreduce, str):
                if allreduce != "nccl":
                    raise ValueError("allreduce must be a callable or 'nccl'")
                # the library's own NCCL communicator (one per handle; the 128-byte id travels over torch.distributed)
                hd = _handle(int(device))
                if getattr(hd, "nccl", None) != (int(rank), int(world)):
                    import torch.distributed as tdist

                    buf = ctypes.create_string_buffer(128)
                    if rank == 0:
                        _lib.check(L.osfm_nccl_unique_id(buf))
                    box = [buf.raw]
                    tdist.broadcast_object_list(box, src=0)
                    idb = ctypes.create_string_buffer(box[0], 128)
                    _lib.check(L.osfm_ba_set_nccl(h, int(rank), int(world), idb))
                    hd.nccl = (int(rank), int(world))
                _lib.check(L.osfm_ba_set_distributed(h, int(rank), int(world), ctypes.cast(None, _lib.ALLREDUCE_FN), None))
            else:
                def _cb(buf, count, strm, user):
                    try:
                        allreduce(int(buf), int(count), int(strm or 0))
                        return 0
                    except Exception:  # surfaces as RuntimeError from run()
                        import traceback

                        traceback.print_exc()
                        return 1

                cb = _lib.ALLREDUCE_FN(_cb)
                _lib.check(L.osfm_ba_set_distributed(h, int(rank), int(world), cb, None))
        else:
            # the handle is reused between calls: reset whatever a previous distributed solve left
            _lib.check(L.osfm_ba_set_distributed(h, 0, 1, ctypes.cast(None, _lib.ALLREDUCE_FN), None))
        _lib.check(L.osfm_ba_set_stream(h, ctypes.c_void_p(stream) if stream is not None else None))
        _lib.check(L.osfm_ba_run(h))
        cam = np.zeros_like(keep[1])
        inst = np.zeros((NI, 6))
        rc = np.zeros((NR, 6))
        def _out(name, shape):
            a = out.get(name) if out else None
            if a is None:
                return np.empty(shape)
            if a.shape != shape or a.dtype != np.float64 or not a.flags.c_contiguous:
                raise ValueError("out[%r] must be a C-contiguous float64 array of shape %r" % (name, shape))
            return a

        pts = _out("points", (P, 3))
        rep = _out("reprojection_errors", (N, 3))
        if not compute_reprojection_errors:
            rep[:] = 0.0
        _lib.check(L.osfm_ba_get_cameras(h, _p(cam)))
        _lib.check(L.osfm_ba_get_rig_instances(h, _p(inst)))
        _lib.check(L.osfm_ba_get_rig_cameras(h, _p(rc)))
        _lib.check(L.osfm_ba_get_points(h, _p(pts)))
        if compute_reprojection_errors:
            _lib.check(L.osfm_ba_get_reprojection_errors(h, _p(rep)))
        s = _lib.BASummary()
        _lib.check(L.osfm_ba_get_summary(h, ctypes.byref(s)))
        summary = {f[0]: getattr(s, f[0]) for f in s._fields_}
        summary["message"] = s.message.decode()
        summary["termination"] = _TERMINATION[s.termination]
        return {"cam_params": cam, "inst": inst, "rigcam": rc, "points": pts, "reprojection_errors": rep,
                "summary": summary}


def eval_observation(projection_type: int, camera, rig_instance, rig_camera, use_rig_camera: bool, point, observed,
                     std_deviation: float, device: int = 0):
    """Residual and Jacobian blocks of one observation computed on the GPU (test hook)."""
    L = _lib.load()
    C = bp.camera_num_params(projection_type)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    cam, ri, pt, ob = f64(camera), f64(rig_instance), f64(point), f64(observed)
    rc = f64(rig_camera if rig_camera is not None else np.zeros(6))
    r, jc, ji, jrc, jp = np.zeros(3), np.zeros(3 * 16), np.zeros(18), np.zeros(18), np.zeros(9)
    n = ctypes.c_int()
    _lib.check(L.osfm_ba_eval_observation(device, int(projection_type), _p(cam), _p(ri), _p(rc), int(bool(use_rig_camera)),
                                          _p(pt), _p(ob), float(std_deviation), _p(r), _p(jc), _p(ji), _p(jrc), _p(jp),
                                          ctypes.byref(n)))
    k = n.value
    return (r[:k].copy(), jc[:k * C].reshape(k, C).copy(), ji[:k * 6].reshape(k, 6).copy(),
            jrc[:k * 6].reshape(k, 6).copy(), jp[:k * 3].reshape(k, 3).copy())


class Point:
    """bundle::Point as exposed by pybundle (pybind.cc:34-43)."""

    def __init__(self, pid: str, p: np.ndarray):
        self.id = pid
        self.p = p
        self.reprojection_errors: Dict[str, np.ndarray] = {}


def _key(s) -> str:
    # ids may be bytes or unicode (opensfm/test/test_bundle.py:20-34)
    return s.decode("utf-8", "replace") if isinstance(s, bytes) else str(s)


class BundleAdjuster:
    """`pybundle.BundleAdjuster` backed by the GPU engine."""

    def __init__(self, device: int = 0):
        self.device = device
        self._cams: Dict[str, Dict[str, Any]] = {}
        self._rig_cameras: Dict[str, Dict[str, Any]] = {}
        self._instances: Dict[str, Dict[str, Any]] = {}
        self._shots: Dict[str, Dict[str, str]] = {}
        self._points: Dict[str, Dict[str, Any]] = {}
        self._obs: List[Any] = []
        # defaults of bundle::BundleAdjuster() (bundle_adjuster.cc:24-44)
        self._loss = ("CauchyLoss", 1.0)
        self._prior_sd = dict(focal_sd=1.0, aspect_ratio_sd=1.0, c_sd=1.0, k1_sd=1.0, k2_sd=1.0, p1_sd=1.0,
                              p2_sd=1.0, k3_sd=1.0, k4_sd=1.0)
        self._max_iterations = 500
        self._num_threads = 1
        self._linear_solver = "SPARSE_SCHUR"
        self._compute_reprojection_errors = True
        self._use_analytic = False
        self._summary: Optional[Dict[str, Any]] = None

    # -- cameras ---------------------------------------------------------
    def add_camera(self, cid, camera, prior, constant: bool) -> None:
        self._cams[_key(cid)] = dict(type=T.camera_type_id(camera), values=T.camera_values(camera),
                                     prior=T.camera_values(prior), constant=bool(constant), proto=camera)

    def get_camera(self, cid):
        c = self._cams.get(_key(cid))
        if c is None:
            raise RuntimeError("Camera %s doesn't exist." % _key(cid))  # bundle_adjuster.cc:104
        names = {v: k for k, v in bp.PROJECTION_NAMES.items() if k != "equirectangular"}
        out = T.Camera(names[c["type"]], c["values"])
        out.id = _key(cid)
        return out

    # -- rig cameras / instances ------------------------------------------
    def add_rig_camera(self, rid, pose, prior_pose, fixed: bool) -> None:
        self._rig_cameras[_key(rid)] = dict(params=T.pose_to_ba_params(pose), prior=T.pose_to_ba_params(prior_pose),
                                            constant=bool(fixed))

    def get_rig_camera_pose(self, rid):
        r = self._rig_cameras.get(_key(rid))
        if r is None:
            raise RuntimeError("Rig camera %s doesn't exist." % _key(rid))
        return T.Pose.from_ba_params(r["params"])

    def add_rig_instance(self, iid, pose, shot_cameras: Dict[str, str], shot_rig_cameras: Dict[str, str],
                         fixed: bool) -> None:
        iid = _key(iid)
        for shot_id, cam_id in shot_cameras.items():
            if _key(cam_id) not in self._cams:
                raise RuntimeError("Camera %s doesn't exist." % _key(cam_id))  # bundle_adjuster.cc:130
            rc_id = _key(shot_rig_cameras[shot_id])
            if rc_id not in self._rig_cameras:
                raise RuntimeError("Rig camera %s doesn't exist." % rc_id)
            self._shots[_key(shot_id)] = dict(instance=iid, camera=_key(cam_id), rig_camera=rc_id)
        self._instances[iid] = dict(params=T.pose_to_ba_params(pose), constant=bool(fixed), prior=None)

    def get_rig_instance_pose(self, iid):
        r = self._instances.get(_key(iid))
        if r is None:
            raise RuntimeError("Rig instance %s doesn't exist." % _key(iid))
        return T.Pose.from_ba_params(r["params"])

    def add_rig_instance_position_prior(self, iid, position, std_deviation, scale_group: str = "") -> None:
        r = self._instances.get(_key(iid))
        if r is None:
            raise RuntimeError("Rig instance %s doesn't exist." % _key(iid))  # bundle_adjuster.cc:170
        r["prior"] = (np.asarray(position, dtype=np.float64), np.asarray(std_deviation, dtype=np.float64))

    # -- points / observations --------------------------------------------
    def add_point(self, pid, position, constant: bool) -> None:
        self._points[_key(pid)] = dict(p=np.asarray(position, dtype=np.float64).copy(), constant=bool(constant),
                                       errors={})

    def has_point(self, pid) -> bool:
        return _key(pid) in self._points

    def get_point(self, pid) -> Point:
        r = self._points.get(_key(pid))
        if r is None:
            raise RuntimeError("Point %s doesn't exist." % _key(pid))
        pt = Point(_key(pid), r["p"].copy())
        pt.reprojection_errors = dict(r["errors"])
        return pt

    def add_point_projection_observation(self, shot, point, observation, std_deviation, depth_prior=None) -> None:
        shot, point = _key(shot), _key(point)
        if shot not in self._shots or point not in self._points:
            # the reference uses std::map::at (bundle_adjuster.cc:242-244) -> IndexError in Python
            raise IndexError("map::at")
        if depth_prior is not None:
            raise NotImplementedError("relative depth priors are outside this engine's scope (SURVEY.md §8a)")
        self._obs.append((shot, point, float(observation[0]), float(observation[1]), float(std_deviation)))

    def add_observations_bulk(self, shots: Sequence[str], points: Sequence[str], xy: np.ndarray,
                              std_deviation: np.ndarray) -> None:
        """Bulk form of add_point_projection_observation (SURVEY.md §7 'String-keyed API')."""
        for s, p, o, sd in zip(shots, points, np.asarray(xy), np.asarray(std_deviation)):
            self.add_point_projection_observation(s, p, o, sd)

    # -- options -----------------------------------------------------------
    def set_point_projection_loss_function(self, name: str, threshold: float) -> None:
        self._loss = (name, float(threshold))

    def set_relative_motion_loss_function(self, name: str, threshold: float) -> None:
        pass  # no relative-motion residuals in this engine

    def set_internal_parameters_prior_sd(self, focal_sd, aspect_ratio_sd, c_sd, k1_sd, k2_sd, p1_sd, p2_sd, k3_sd,
                                         k4_sd) -> None:
        self._prior_sd = dict(focal_sd=focal_sd, aspect_ratio_sd=aspect_ratio_sd, c_sd=c_sd, k1_sd=k1_sd,
                              k2_sd=k2_sd, p1_sd=p1_sd, p2_sd=p2_sd, k3_sd=k3_sd, k4_sd=k4_sd)

    def set_max_num_iterations(self, n: int) -> None:
        self._max_iterations = int(n)

    def set_num_threads(self, n: int) -> None:
        self._num_threads = int(n)

    def set_use_analytic_derivatives(self, use: bool) -> None:
        self._use_analytic = bool(use)  # the GPU engine is always analytic

    def set_linear_solver_type(self, name: str) -> None:
        self._linear_solver = name

    def set_compute_reprojection_errors(self, v: bool) -> None:
        self._compute_reprojection_errors = bool(v)

    def set_compute_covariances(self, v: bool) -> None:
        if v:
            raise NotImplementedError("covariance estimation is outside this engine's scope")

    def get_covariance_estimation_valid(self) -> bool:
        return False

    def set_adjust_absolute_position_std(self, v: bool) -> None:
        if v:
            raise NotImplementedError("adjust_absolute_position_std is outside this engine's scope")

    def _unsupported(self, *a, **k):
        raise NotImplementedError("this residual type is outside the hot path this engine replaces (SURVEY.md §8a)")

    add_point_prior = add_reconstruction = add_reconstruction_instance = set_scale_sharing = _unsupported
    add_relative_motion = add_relative_rotation = add_common_position = add_heatmap = _unsupported
    add_absolute_position_heatmap = add_absolute_up_vector = add_absolute_pan = add_absolute_tilt = _unsupported
    add_absolute_roll = add_linear_motion = set_gauge_fix_shots = get_reconstruction = _unsupported

    # -- run ---------------------------------------------------------------
    def to_problem(self) -> bp.BAProblem:
        cam_ids = list(self._cams)
        cam_index = {c: i for i, c in enumerate(cam_ids)}
        inst_ids = list(self._instances)
        inst_index = {c: i for i, c in enumerate(inst_ids)}
        rc_ids = list(self._rig_cameras)
        rc_index = {c: i for i, c in enumerate(rc_ids)}
        shot_ids = list(self._shots)
        shot_index = {c: i for i, c in enumerate(shot_ids)}
        pt_ids = list(self._points)
        pt_index = {c: i for i, c in enumerate(pt_ids)}
        self._order = (cam_ids, inst_ids, rc_ids, shot_ids, pt_ids)
        NI = len(inst_ids)
        rigcam = np.array([self._rig_cameras[r]["params"] for r in rc_ids]).reshape(-1, 6) if rc_ids else np.zeros((1, 6))
        rc_const = np.array([self._rig_cameras[r]["constant"] for r in rc_ids], dtype=np.int32) if rc_ids else np.ones(1, dtype=np.int32)
        # IsRigCameraUseful (bundle_adjuster.cc:17-20): free parameters or a non-zero pose
        useful = {r: (not self._rig_cameras[r]["constant"]) or bool(np.any(self._rig_cameras[r]["params"] != 0.0))
                  for r in rc_ids}
        has_prior = np.zeros(NI, dtype=np.int32)
        ppos = np.zeros((NI, 3))
        pstd = np.ones((NI, 3))
        for i, iid in enumerate(inst_ids):
            pr = self._instances[iid]["prior"]
            if pr is not None:
                has_prior[i] = 1
                ppos[i], pstd[i] = pr
        n_obs = len(self._obs)
        obs_shot = np.fromiter((shot_index[o[0]] for o in self._obs), dtype=np.int32, count=n_obs)
        obs_point = np.fromiter((pt_index[o[1]] for o in self._obs), dtype=np.int32, count=n_obs)
        obs_xy = np.array([[o[2], o[3]] for o in self._obs], dtype=np.float64).reshape(-1, 2)
        obs_sigma = np.array([o[4] for o in self._obs], dtype=np.float64)
        pb = bp.make_problem(
            [self._cams[c]["type"] for c in cam_ids], [self._cams[c]["values"] for c in cam_ids],
            np.array([self._instances[i]["params"] for i in inst_ids]).reshape(-1, 6),
            np.array([self._points[p]["p"] for p in pt_ids]).reshape(-1, 3),
            obs_shot, obs_point, obs_xy, obs_sigma,
            shot_inst=[inst_index[self._shots[s]["instance"]] for s in shot_ids],
            shot_cam=[cam_index[self._shots[s]["camera"]] for s in shot_ids],
            rigcam=rigcam, shot_rc=[rc_index[self._shots[s]["rig_camera"]] for s in shot_ids],
            shot_use_rc=[int(useful[self._shots[s]["rig_camera"]]) for s in shot_ids],
            cam_const=[int(self._cams[c]["constant"]) for c in cam_ids],
            inst_const=[int(self._instances[i]["constant"]) for i in inst_ids],
            rigcam_const=rc_const, point_const=[int(self._points[p]["constant"]) for p in pt_ids],
            cam_prior_list=[self._cams[c]["prior"] for c in cam_ids], prior_sd=self._prior_sd,
            loss_name=self._loss[0], loss_threshold=self._loss[1], max_iterations=self._max_iterations,
            linear_solver_type=self._linear_solver, num_threads=self._num_threads)
        pb.inst_has_prior = has_prior
        pb.inst_prior_pos = ppos
        pb.inst_prior_std = pstd
        return pb

    def run(self) -> None:
        pb = self.to_problem()
        res = solve(pb, device=self.device, compute_reprojection_errors=self._compute_reprojection_errors)
        cam_ids, inst_ids, rc_ids, shot_ids, pt_ids = self._order
        off = pb.cam_off
        for i, c in enumerate(cam_ids):
            self._cams[c]["values"] = res["cam_params"][off[i]:off[i + 1]].copy()
        for i, iid in enumerate(inst_ids):
            self._instances[iid]["params"] = res["inst"][i].copy()
        for i, r in enumerate(rc_ids):
            self._rig_cameras[r]["params"] = res["rigcam"][i].copy()
        for i, p in enumerate(pt_ids):
            self._points[p]["p"] = res["points"][i].copy()
            self._points[p]["errors"] = {}
        if self._compute_reprojection_errors:
            rep = res["reprojection_errors"]
            spherical = {s: self._cams[self._shots[s]["camera"]]["type"] == bp.SPHERICAL for s in shot_ids}
            for k, o in enumerate(self._obs):
                self._points[o[1]]["errors"][o[0]] = rep[k, :3].copy() if spherical[o[0]] else rep[k, :2].copy()
        self._summary = res["summary"]

    def brief_report(self) -> str:
        s = self._summary
        if s is None:
            return "Solver has not run."
        return ("opensfm_b200 BA Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s"
                % (s["iterations"], s["initial_cost"], s["final_cost"], s["termination"]))

    def full_report(self) -> str:
        s = self._summary
        if s is None:
            return "Solver has not run."
        return self.brief_report() + "\n" + "\n".join("%s: %s" % kv for kv in s.items())
