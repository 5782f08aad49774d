"""Minimal host-side stand-ins for `pygeometry.Camera` / `pygeometry.Pose`.

The reference's BundleAdjuster API takes these pybind objects
(opensfm/src/bundle/python/pybind.cc:52-66).  They cannot be built in this image
(Eigen missing), so the boundary accepts any duck-typed object with the same
attribute names; these two classes provide them for tests, fixtures and users
without the reference installed.  Pure data + fp64 numpy; nothing here is on the
hot path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import ba_problem as bp


def _rotvec_to_matrix(r: np.ndarray) -> np.ndarray:
    from scipy.spatial.transform import Rotation

    return Rotation.from_rotvec(np.asarray(r, dtype=np.float64)).as_matrix()


def _matrix_to_rotvec(R: np.ndarray) -> np.ndarray:
    from scipy.spatial.transform import Rotation

    return Rotation.from_matrix(np.asarray(R, dtype=np.float64)).as_rotvec()


class Pose:
    """World-to-camera pose: x_cam = R(rotation) x_world + translation
    (opensfm/src/geometry/pose.h; Python signature Pose(rotation, translation))."""

    def __init__(self, rotation: Optional[Sequence[float]] = None, translation: Optional[Sequence[float]] = None):
        self.rotation = np.zeros(3) if rotation is None else np.asarray(rotation, dtype=np.float64).copy()
        self.translation = np.zeros(3) if translation is None else np.asarray(translation, dtype=np.float64).copy()

    def get_rotation_matrix(self) -> np.ndarray:
        return _rotvec_to_matrix(self.rotation)

    def set_rotation_matrix(self, R: np.ndarray) -> None:
        self.rotation = _matrix_to_rotvec(R)

    def get_origin(self) -> np.ndarray:
        return -self.get_rotation_matrix().T @ self.translation

    def set_origin(self, origin: Sequence[float]) -> None:
        self.translation = -self.get_rotation_matrix() @ np.asarray(origin, dtype=np.float64)

    def get_R_cam_to_world_min(self) -> np.ndarray:
        return -self.rotation

    def to_ba_params(self) -> np.ndarray:
        """[angle-axis camera->world | origin] (bundle/data/pose.h:34-43)."""
        return np.concatenate([-self.rotation, self.get_origin()])

    @staticmethod
    def from_ba_params(p: Sequence[float]) -> "Pose":
        p = np.asarray(p, dtype=np.float64)
        pose = Pose(-p[:3], np.zeros(3))
        pose.set_origin(p[3:6])
        return pose


class Camera:
    """Projection type + parameter vector in the reference's storage order
    (opensfm/src/geometry/src/camera.cc:9-178)."""

    def __init__(self, projection_type: str, values: Sequence[float]):
        self.projection_type = projection_type
        self.type_id = bp.PROJECTION_NAMES[projection_type]
        self.values = np.asarray(values, dtype=np.float64).copy()
        assert len(self.values) == bp.camera_num_params(self.type_id)
        self.id = ""
        self.width = 0
        self.height = 0

    # factory names follow pygeometry.Camera.create_*
    @staticmethod
    def create_perspective(focal, k1, k2):
        return Camera("perspective", [k1, k2, focal])

    @staticmethod
    def create_brown(focal, aspect_ratio, principal_point, distortion):
        return Camera("brown", list(distortion) + [focal, aspect_ratio, principal_point[0], principal_point[1]])

    @staticmethod
    def create_fisheye(focal, k1, k2):
        return Camera("fisheye", [k1, k2, focal])

    @staticmethod
    def create_fisheye_opencv(focal, aspect_ratio, principal_point, distortion):
        return Camera("fisheye_opencv", list(distortion) + [focal, aspect_ratio, principal_point[0], principal_point[1]])

    @staticmethod
    def create_fisheye62(focal, aspect_ratio, principal_point, distortion):
        return Camera("fisheye62", list(distortion) + [focal, aspect_ratio, principal_point[0], principal_point[1]])

    @staticmethod
    def create_fisheye624(focal, aspect_ratio, principal_point, distortion):
        return Camera("fisheye624", list(distortion) + [focal, aspect_ratio, principal_point[0], principal_point[1]])

    @staticmethod
    def create_dual(transition, focal, k1, k2):
        return Camera("dual", [transition, k1, k2, focal])

    @staticmethod
    def create_spherical():
        return Camera("spherical", [0.0])

    @staticmethod
    def create_radial(focal, aspect_ratio, principal_point, distortion):
        return Camera("radial", list(distortion) + [focal, aspect_ratio, principal_point[0], principal_point[1]])

    @staticmethod
    def create_simple_radial(focal, aspect_ratio, principal_point, k1):
        return Camera("simple_radial", [k1, focal, aspect_ratio, principal_point[0], principal_point[1]])

    def get_parameters_values(self) -> np.ndarray:
        return self.values.copy()

    def set_parameters_values(self, v) -> None:
        self.values = np.asarray(v, dtype=np.float64).copy()

    def get_parameters_types(self) -> List[str]:
        return list(bp.CAMERA_PARAM_NAMES[self.type_id])

    def get_parameters_map(self) -> Dict[str, float]:
        return dict(zip(self.get_parameters_types(), self.values.tolist()))

    def __getattr__(self, name):
        # camera.focal, camera.k1, ... like pygeometry.Camera's properties
        if name in ("values", "type_id"):
            raise AttributeError(name)
        names = bp.CAMERA_PARAM_NAMES.get(self.__dict__.get("type_id", -1), [])
        if name in names:
            return float(self.values[names.index(name)])
        raise AttributeError(name)

    def copy(self) -> "Camera":
        c = Camera(self.projection_type, self.values)
        c.id, c.width, c.height = self.id, self.width, self.height
        return c


def camera_values(cam) -> np.ndarray:
    """Parameter vector of a duck-typed camera (ours or pygeometry.Camera)."""
    if hasattr(cam, "get_parameters_values"):
        return np.asarray(cam.get_parameters_values(), dtype=np.float64)
    raise TypeError("camera object must provide get_parameters_values()")


def camera_type_id(cam) -> int:
    pt = cam.projection_type
    if not isinstance(pt, str):
        pt = str(pt).split(".")[-1].lower()
    return bp.PROJECTION_NAMES[pt]


def pose_to_ba_params(pose) -> np.ndarray:
    if hasattr(pose, "to_ba_params"):
        return pose.to_ba_params()
    # pygeometry.Pose: rotation (world->cam angle-axis), get_origin()
    return np.concatenate([-np.asarray(pose.rotation, dtype=np.float64), np.asarray(pose.get_origin(), dtype=np.float64)])
