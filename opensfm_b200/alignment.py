"""`pybundle.ReconstructionAlignment` on the GPU engine  --  the global alignment step of the large pipeline
(opensfm/large/tools.py:278-307; class in opensfm/src/bundle/reconstruction_alignment.h:369-580, bound in
bundle/python/pybind.cc:122-182).

Shots are [R | t] world-to-camera (6 parameters), reconstructions [R | t | scale] (7, scale >= 0.1); the five
constraint kinds are side terms of the engine (include/opensfm_b200.h, OSFM_SIDE_RA_*): the problem has no
points, so the "reduced camera system" is the whole normal matrix and the solve is LM + PCG on it.  Same method
names, argument order and loss functions as the reference (CauchyLoss(1) on relative motions and common cameras,
SoftLOneLoss(1) on relative-absolute positions and common points, none on absolute positions)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from . import ba_problem as bp
from . import bundle as _bundle


class RAShot:
    def __init__(self):
        self.id = ""
        self.parameters = np.zeros(6)
        self.constant = False

    rx = property(lambda s: s.parameters[0]); ry = property(lambda s: s.parameters[1]); rz = property(lambda s: s.parameters[2])
    tx = property(lambda s: s.parameters[3]); ty = property(lambda s: s.parameters[4]); tz = property(lambda s: s.parameters[5])


class RAReconstruction:
    def __init__(self):
        self.id = ""
        self.parameters = np.array([0.0, 0, 0, 0, 0, 0, 1.0])
        self.constant = False

    rx = property(lambda s: s.parameters[0]); ry = property(lambda s: s.parameters[1]); rz = property(lambda s: s.parameters[2])
    tx = property(lambda s: s.parameters[3]); ty = property(lambda s: s.parameters[4]); tz = property(lambda s: s.parameters[5])
    scale = property(lambda s: s.parameters[6])


class RARelativeMotionConstraint:
    def __init__(self, reconstruction, shot, rx, ry, rz, tx, ty, tz):
        self.reconstruction = reconstruction
        self.shot = shot
        self.parameters = np.array([rx, ry, rz, tx, ty, tz], dtype=np.float64)
        self.scale_matrix = np.eye(6)

    def set_scale_matrix(self, i: int, j: int, value: float) -> None:
        self.scale_matrix[i, j] = value


class ReconstructionAlignment:
    def __init__(self, device: int = 0):
        self.device = device
        self._shots: Dict[str, RAShot] = {}
        self._recs: Dict[str, RAReconstruction] = {}
        self._relative_motions: List[RARelativeMotionConstraint] = []
        self._absolute: List[Any] = []
        self._relative_absolute: List[Any] = []
        self._common_cameras: List[Any] = []
        self._common_points: List[Any] = []
        self._summary: Optional[Dict[str, Any]] = None

    def get_shot(self, sid) -> RAShot:
        return self._shots[_bundle._key(sid)]

    def get_reconstruction(self, rid) -> RAReconstruction:
        return self._recs[_bundle._key(rid)]

    def add_shot(self, sid, rx, ry, rz, tx, ty, tz, constant: bool) -> None:
        s = RAShot()
        s.id = _bundle._key(sid)
        s.parameters = np.array([rx, ry, rz, tx, ty, tz], dtype=np.float64)
        s.constant = bool(constant)
        self._shots[s.id] = s

    def add_reconstruction(self, rid, rx, ry, rz, tx, ty, tz, scale, constant: bool) -> None:
        r = RAReconstruction()
        r.id = _bundle._key(rid)
        r.parameters = np.array([rx, ry, rz, tx, ty, tz, scale], dtype=np.float64)
        r.constant = bool(constant)
        self._recs[r.id] = r

    def add_relative_motion_constraint(self, rm: RARelativeMotionConstraint) -> None:
        self._relative_motions.append(rm)

    def add_absolute_position_constraint(self, shot_id, x, y, z, std_deviation) -> None:
        self._absolute.append((_bundle._key(shot_id), np.array([x, y, z], dtype=np.float64), float(std_deviation)))

    def add_relative_absolute_position_constraint(self, reconstruction_id, shot_id, x, y, z, std_deviation) -> None:
        self._relative_absolute.append((_bundle._key(reconstruction_id), _bundle._key(shot_id),
                                        np.array([x, y, z], dtype=np.float64), float(std_deviation)))

    def add_common_point_constraint(self, reconstruction_a_id, xa, ya, za, reconstruction_b_id, xb, yb, zb,
                                    std_deviation) -> None:
        self._common_points.append((_bundle._key(reconstruction_a_id), np.array([xa, ya, za], dtype=np.float64),
                                    _bundle._key(reconstruction_b_id), np.array([xb, yb, zb], dtype=np.float64),
                                    float(std_deviation)))

    def add_common_camera_constraint(self, reconstruction_a_id, shot_a_id, reconstruction_b_id, shot_b_id,
                                     std_deviation_center, std_deviation_rotation) -> None:
        self._common_cameras.append((_bundle._key(reconstruction_a_id), _bundle._key(shot_a_id),
                                     _bundle._key(reconstruction_b_id), _bundle._key(shot_b_id),
                                     float(std_deviation_center), float(std_deviation_rotation)))

    # -- the problem -------------------------------------------------------------------------------------------
    def to_problem(self) -> bp.BAProblem:
        shot_ids, rec_ids = list(self._shots), list(self._recs)
        si = {s: i for i, s in enumerate(shot_ids)}
        NI = len(shot_ids)
        z_i = lambda n: np.zeros(n, dtype=np.int32)
        pb = bp.BAProblem(
            cam_type=z_i(0), cam_params=np.zeros(0), cam_const=z_i(0), cam_prior=np.zeros(0), cam_prior_sigma=np.zeros(0),
            cam_prior_log=z_i(0),
            inst=np.array([self._shots[s].parameters for s in shot_ids], dtype=np.float64).reshape(-1, 6),
            inst_const=np.array([int(self._shots[s].constant) for s in shot_ids], dtype=np.int32),
            inst_has_prior=z_i(NI), inst_prior_pos=np.zeros((NI, 3)), inst_prior_std=np.ones((NI, 3)),
            rigcam=np.zeros((1, 6)), rigcam_const=np.ones(1, dtype=np.int32),
            shot_inst=z_i(0), shot_cam=z_i(0), shot_rc=z_i(0), shot_use_rc=z_i(0),
            points=np.zeros((0, 3)), point_const=z_i(0), obs_shot=z_i(0), obs_point=z_i(0), obs_xy=np.zeros((0, 2)),
            obs_sigma=np.zeros(0), loss_name="TrivialLoss", loss_threshold=1.0, max_iterations=500,
            linear_solver_type="SPARSE_NORMAL_CHOLESKY", num_threads=8)
        # reconstructions: 7-parameter ext blocks, scale bounded below by 0.1 when free (:478-491)
        ri = {}
        for r in rec_ids:
            rec = self._recs[r]
            ri[r] = pb.add_ext_block(rec.parameters, rec.constant, [-np.inf] * 6 + [0.1])
        self._order = (shot_ids, rec_ids, ri)
        T = pb.side_terms
        for rm in self._relative_motions:
            T.append(bp.SideTerm(bp.SIDE_RA_RELATIVE_MOTION, 6, [(bp.SB_EXT, ri[_bundle._key(rm.reconstruction)]),
                                                                  (bp.SB_INST, si[_bundle._key(rm.shot)])],
                                 np.concatenate([rm.parameters, rm.scale_matrix.reshape(36)]), bp.LOSS_CAUCHY, 1.0))
        for sid, pos, sd in self._absolute:
            T.append(bp.SideTerm(bp.SIDE_RA_ABSOLUTE_POSITION, 3, [(bp.SB_INST, si[sid])], np.concatenate([pos, [1.0 / sd]])))
        for rid, sid, pos, sd in self._relative_absolute:
            T.append(bp.SideTerm(bp.SIDE_RA_RELATIVE_ABSOLUTE_POSITION, 3, [(bp.SB_EXT, ri[rid])],
                                 np.concatenate([pos, self._shots[sid].parameters, [1.0 / sd]]), bp.LOSS_SOFTLONE, 1.0))
        for ra, sa, rb, sb, sdc, sdr in self._common_cameras:
            T.append(bp.SideTerm(bp.SIDE_RA_COMMON_CAMERA, 6, [(bp.SB_EXT, ri[ra]), (bp.SB_EXT, ri[rb])],
                                 np.concatenate([self._shots[sa].parameters, self._shots[sb].parameters, [1.0 / sdc, 1.0 / sdr]]),
                                 bp.LOSS_CAUCHY, 1.0))
        for ra, pa, rb, pbt, sd in self._common_points:
            T.append(bp.SideTerm(bp.SIDE_RA_COMMON_POINT, 3, [(bp.SB_EXT, ri[ra]), (bp.SB_EXT, ri[rb])],
                                 np.concatenate([pa, pbt, [1.0 / sd]]), bp.LOSS_SOFTLONE, 1.0))
        return pb

    def apply_results(self, pb: bp.BAProblem, res: Dict[str, Any]) -> None:
        shot_ids, rec_ids, ri = self._order
        for i, s in enumerate(shot_ids):
            self._shots[s].parameters = res["inst"][i].copy()
        eo = pb.ext_off
        for r in rec_ids:
            self._recs[r].parameters = res["ext_values"][eo[ri[r]]:eo[ri[r] + 1]].copy()
        self._summary = res["summary"]

    def run(self) -> None:
        pb = self.to_problem()
        self.apply_results(pb, _bundle.solve(pb, device=self.device, compute_reprojection_errors=False))

    def brief_report(self) -> str:
        s = self._summary
        if s is None:
            return "Solver has not run."
        return ("opensfm_b200 alignment Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s"
                % (s["iterations"], s["initial_cost"], s["final_cost"], s["termination"]))

    def full_report(self) -> str:
        return self.brief_report()
