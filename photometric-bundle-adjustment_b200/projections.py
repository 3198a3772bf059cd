"""Host-side mirror of the SfM-loop callers around bundle_adjustment()
(SURVEY.md §8(f)-2/3): `Landmark::get_p` (include/visnav/common_types.h:205-217),
`compute_projections` + `set_outlier_flags` (src/sfm.cpp:1928-2008) and the
keep/remove decision of `remove_outlier_landmarks` (src/sfm.cpp:2029-2100).
All compute is the CUDA path behind include/pba.h; there is no CPU fallback.
"""
import ctypes as C
import dataclasses

import numpy as np

from . import _ffi
from .problem import Problem

# OutlierFlags (common_types.h:277-285)
OutlierNone = 0
OutlierReprojectionErrorHuge = 1 << 0
OutlierReprojectionErrorNormal = 1 << 1
OutlierCameraDistance = 1 << 2
OutlierZCoordinate = 1 << 3


@dataclasses.dataclass
class ProjectionThresholds:
    """The pangolin::Var thresholds of src/sfm.cpp:254-261, same defaults."""
    reprojection_error_huge_pixel: float = 40.0
    reprojection_error_normal_pixel: float = 3.0
    camera_center_distance_meter: float = 0.1
    z_coordinate_meter: float = 0.05

    def to_c(self):
        t = _ffi.pba_projection_thresholds()
        for f in dataclasses.fields(self):
            setattr(t, f.name, float(getattr(self, f.name)))
        return t


@dataclasses.dataclass
class Projections:
    """ProjectedLandmark fields (common_types.h:288-297) as flat arrays over the
    observation slots (host observation of a landmark first, then its others)."""
    slot_ptr: np.ndarray            # [n_landmarks+1] first slot of each landmark
    point_reprojected: np.ndarray   # [n_slots, 2]
    point_3d_c: np.ndarray          # [n_slots, 3]
    reprojection_error: np.ndarray  # [n_slots]
    outlier_flags: np.ndarray       # [n_slots] uint32
    landmark_remove: np.ndarray     # [n_landmarks] uint8
    any_severe_outliers: bool

    def is_landmark_outlier(self, l):
        """src/sfm.cpp:2012-2020."""
        return bool(np.any(self.outlier_flags[self.slot_ptr[l]:self.slot_ptr[l + 1]] != OutlierNone))


def landmark_positions(problem: Problem, device=0):
    """Landmark::get_p for every landmark -> [n_landmarks, 3] world points."""
    out = np.zeros((problem.n_landmarks, 3))
    pc = problem.c
    _ffi.check(_ffi.load_lib().pba_landmark_positions(C.byref(pc), int(device), _ffi.ptr(out, C.c_double)),
               "pba_landmark_positions")
    return out


def compute_projections(problem: Problem, thresholds: ProjectionThresholds = None, device=0) -> Projections:
    thresholds = thresholds or ProjectionThresholds()
    nl = problem.n_landmarks
    ns = int(problem.n_obs) + nl
    repro, p3c, err = np.zeros((ns, 2)), np.zeros((ns, 3)), np.zeros(ns)
    flags, remove = np.zeros(ns, np.uint32), np.zeros(nl, np.uint8)
    severe = C.c_int32(0)
    t = thresholds.to_c()
    pc = problem.c
    _ffi.check(_ffi.load_lib().pba_compute_projections(
        C.byref(pc), C.byref(t), int(device), _ffi.ptr(repro, C.c_double), _ffi.ptr(p3c, C.c_double),
        _ffi.ptr(err, C.c_double), _ffi.ptr(flags, C.c_uint32), _ffi.ptr(remove, C.c_uint8), C.byref(severe)),
        "pba_compute_projections")
    slot_ptr = problem.lm_obs_ptr + np.arange(nl + 1, dtype=np.int64)
    return Projections(slot_ptr, repro, p3c, err, flags, remove, bool(severe.value))


def triangulate_inverse_depth(model0, intr0, model1, intr1, T_w_c0, T_w_c1, uv0, uv1, device=0):
    """add_new_landmarks_between_cams (include/visnav/map_utils.h:121-195) for n shared tracks: returns
    (p in camera 0's frame [n,3], initial inverse distance [n])."""
    m0 = _ffi.CAM_NAMES[model0] if isinstance(model0, str) else int(model0)
    m1 = _ffi.CAM_NAMES[model1] if isinstance(model1, str) else int(model1)
    i0, i1 = np.ascontiguousarray(intr0, np.float64), np.ascontiguousarray(intr1, np.float64)
    T0, T1 = np.ascontiguousarray(T_w_c0, np.float64), np.ascontiguousarray(T_w_c1, np.float64)
    uv0, uv1 = np.ascontiguousarray(uv0, np.float64).reshape(-1, 2), np.ascontiguousarray(uv1, np.float64).reshape(-1, 2)
    n = uv0.shape[0]
    p = np.zeros((n, 3))
    rho = np.zeros(n)
    _ffi.check(_ffi.load_lib().pba_triangulate_inverse_depth(
        m0, _ffi.ptr(i0, C.c_double), m1, _ffi.ptr(i1, C.c_double), _ffi.ptr(T0, C.c_double), _ffi.ptr(T1, C.c_double), n,
        _ffi.ptr(uv0, C.c_double), _ffi.ptr(uv1, C.c_double), device, _ffi.ptr(p, C.c_double), _ffi.ptr(rho, C.c_double)),
        "pba_triangulate_inverse_depth")
    return p, rho
