"""ctypes binding of the C ABI declared in include/opensfm_b200.h.

There is no CPU fallback: if the shared library is missing or no CUDA device is
present, calls raise.  (The library is built in-tree by `__graft_entry__.build()`.)
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libopensfm_b200.so")

OSFM_OK = 0
PROJECTION_TYPES = dict(PERSPECTIVE=0, BROWN=1, FISHEYE=2, FISHEYE_OPENCV=3, FISHEYE62=4, FISHEYE624=5,
                        SPHERICAL=6, DUAL=7, RADIAL=8, SIMPLE_RADIAL=9)
LOSS_IDS = {"TrivialLoss": 0, "HuberLoss": 1, "SoftLOneLoss": 2, "CauchyLoss": 3, "ArctanLoss": 4}
LOSS_TUKEY = 5  # side terms only (common position)
SIDE_TYPES = dict(UP_VECTOR=0, PAN=1, TILT=2, ROLL=3, RELATIVE_MOTION=4, RELATIVE_ROTATION=5, COMMON_POSITION=6,
                  LINEAR_MOTION=7, TRANSLATION_PRIOR=8, PARAMETER_BARRIER=9, STD_DEVIATION=10, POSITION_PRIOR=11)
SB_CAM, SB_INST, SB_RIGCAM, SB_EXT = 0, 1, 2, 3


class SideTerm(ctypes.Structure):
    """osfm_side_term (include/opensfm_b200.h)."""
    _fields_ = [("type", c_int32), ("nres", c_int32), ("nblocks", c_int32), ("kind", c_int32 * 6), ("idx", c_int32 * 6),
                ("loss", c_int32), ("loss_a", c_double), ("cofs", c_int32), ("aux", c_int32 * 4)]


class BASummary(ctypes.Structure):
    _fields_ = [
        ("iterations", c_int), ("successful_steps", c_int), ("linear_solves", c_int), ("pcg_iterations", c_int),
        ("termination", c_int), ("initial_cost", c_double), ("final_cost", c_double), ("time_run_s", c_double),
        ("time_device_ms", c_double), ("time_linearize_ms", c_double), ("linearize_launches", c_int64),
        ("time_schur_ms", c_double), ("schur_launches", c_int64), ("time_pcg_ms", c_double),
        ("time_backsub_ms", c_double), ("num_observations_local", c_int64), ("reduced_dim", c_int),
        ("reduced_blocks", c_int), ("reduced_nnz", c_int64), ("jac_planes", c_int), ("kernel_launches", c_int64), ("message", ctypes.c_char * 128),
    ]


ALLREDUCE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int64, c_void_p, c_void_p)

# name -> (restype, argtypes); every symbol include/opensfm_b200.h declares
SIGNATURES = {
    "osfm_last_error": (c_char_p, []),
    "osfm_version": (c_int, []),
    "osfm_kernel_launch_count": (c_int64, []),
    "osfm_matcher_create": (c_int, [c_int, POINTER(c_void_p)]),
    "osfm_matcher_destroy": (c_int, [c_void_p]),
    "osfm_bf_match_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_double, c_void_p, c_int, c_void_p]),
    "osfm_bf_match_u8": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_double, c_void_p, c_int, c_void_p]),
    "osfm_matcher_add_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int)]),
    "osfm_matcher_add_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int)]),
    "osfm_matcher_add_batch_f32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "osfm_matcher_add_batch_u8": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "osfm_matcher_add_u8_l2": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int)]),
    "osfm_matcher_add_batch_u8_l2": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "osfm_matcher_remove": (c_int, [c_void_p, c_int]),
    "osfm_matcher_clear": (c_int, [c_void_p]),
    "osfm_matcher_match_pairs_async": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_double, c_int]),
    "osfm_matcher_set_bearings": (c_int, [c_void_p, c_int, c_void_p]),
    "osfm_matcher_match_pairs_guided_async": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double,
                                                       c_int]),
    "osfm_matcher_sync": (c_int, [c_void_p]),
    "osfm_matcher_fetch": (c_int, [c_void_p, c_void_p, c_int64]),
    "osfm_matcher_fetch_pairs": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_int64)]),
    "osfm_matcher_last_device_ms": (c_int, [c_void_p, POINTER(c_float), POINTER(c_float)]),
    "osfm_matcher_set_kernel": (c_int, [c_void_p, c_int]),
    "osfm_matcher_last_kernel": (c_int, [c_void_p]),
    "osfm_matcher_device_bytes": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "osfm_match_words": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_float,
                                  c_int, c_void_p]),
    "osfm_vlad_distances": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "osfm_ba_create": (c_int, [c_int, POINTER(c_void_p)]),
    "osfm_ba_destroy": (c_int, [c_void_p]),
    "osfm_camera_num_params": (c_int, [c_int]),
    "osfm_ba_set_cameras": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osfm_ba_set_rig_instances": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osfm_ba_set_rig_cameras": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "osfm_ba_set_shots": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osfm_ba_set_points": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "osfm_ba_set_observations": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osfm_ba_set_observations_async": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osfm_ba_set_rig_camera_priors": (c_int, [c_void_p, c_void_p, c_void_p]),
    "osfm_ba_set_point_priors": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osfm_ba_set_ext_blocks": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "osfm_ba_get_ext_blocks": (c_int, [c_void_p, c_void_p]),
    "osfm_ba_set_side_terms": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "osfm_ba_set_options": (c_int, [c_void_p, c_int, c_double, c_int, c_char_p, c_int]),
    "osfm_ba_set_distributed": (c_int, [c_void_p, c_int, c_int, ALLREDUCE_FN, c_void_p]),
    "osfm_nccl_unique_id": (c_int, [c_void_p]),
    "osfm_ba_set_nccl": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "osfm_ba_set_stream": (c_int, [c_void_p, c_void_p]),
    "osfm_ba_run": (c_int, [c_void_p]),
    "osfm_ba_get_summary": (c_int, [c_void_p, POINTER(BASummary)]),
    "osfm_ba_get_cameras": (c_int, [c_void_p, c_void_p]),
    "osfm_ba_get_rig_instances": (c_int, [c_void_p, c_void_p]),
    "osfm_ba_get_rig_cameras": (c_int, [c_void_p, c_void_p]),
    "osfm_ba_get_points": (c_int, [c_void_p, c_void_p]),
    "osfm_ba_get_reprojection_errors": (c_int, [c_void_p, c_void_p]),
    "osfm_ba_eval_observation": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                         c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int)]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library.  Raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "opensfm_b200: %s is missing; run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != OSFM_OK:
        msg = load().osfm_last_error().decode("utf-8", "replace")
        if code == 2:
            raise ValueError(msg)
        raise RuntimeError(msg)
