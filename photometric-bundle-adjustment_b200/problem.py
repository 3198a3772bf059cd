"""Flat (SoA) BA problem on the host: numpy arrays <-> the C ABI's pba_problem.

Mirrors the reference's containers for the BA call (include/visnav/common_types.h:
Cameras = {FrameCamId -> T_w_c}, Landmarks = {TrackId -> inv_depth, obs}, Corners,
Calibration) in the layout include/pba.h documents.
"""
import ctypes as C

import numpy as np

from . import _ffi


class Problem:
    """Owns the host arrays; `.c` is a pba_problem view onto them (no copies)."""

    FIELDS = ("poses", "pose_fixed", "pose_calib", "calib_model", "intrinsics", "inv_depth", "lm_host",
              "lm_host_uv", "lm_obs_ptr", "obs_target", "obs_uv", "images", "affine")

    def __init__(self, mode, poses, pose_fixed, pose_calib, calib_model, intrinsics, inv_depth, lm_host,
                 lm_host_uv, lm_obs_ptr, obs_target, obs_uv=None, images=None, affine=None):
        self.mode = int(mode)
        self.poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        self.pose_fixed = np.ascontiguousarray(pose_fixed, np.uint8).reshape(-1)
        self.pose_calib = np.ascontiguousarray(pose_calib, np.int32).reshape(-1)
        self.calib_model = np.ascontiguousarray(calib_model, np.int32).reshape(-1)
        self.intrinsics = np.ascontiguousarray(intrinsics, np.float64).reshape(-1, 8)
        self.inv_depth = np.ascontiguousarray(inv_depth, np.float64).reshape(-1)
        self.lm_host = np.ascontiguousarray(lm_host, np.int32).reshape(-1)
        self.lm_host_uv = np.ascontiguousarray(lm_host_uv, np.float64).reshape(-1, 2)
        self.lm_obs_ptr = np.ascontiguousarray(lm_obs_ptr, np.int64).reshape(-1)
        self.obs_target = np.ascontiguousarray(obs_target, np.int32).reshape(-1)
        self.obs_uv = None if obs_uv is None else np.ascontiguousarray(obs_uv, np.float64).reshape(-1, 2)
        self.images = None if images is None else np.ascontiguousarray(images, np.uint8)
        self.affine = None if affine is None else np.ascontiguousarray(affine, np.float64).reshape(-1, 2)
        if self.mode == _ffi.MODE_PHOTOMETRIC:
            assert self.images is not None and self.images.ndim == 3, "photometric mode needs images [n,h,w]"
            if self.affine is None:
                self.affine = np.zeros((self.n_poses, 2))
        self._c = None

    n_poses = property(lambda s: s.poses.shape[0])
    n_calib = property(lambda s: s.intrinsics.shape[0])
    n_landmarks = property(lambda s: s.inv_depth.shape[0])
    n_obs = property(lambda s: s.obs_target.shape[0])
    res_per_obs = property(lambda s: 8 if s.mode == _ffi.MODE_PHOTOMETRIC else 2)
    cols_per_obs = property(lambda s: 15 if s.mode == _ffi.MODE_PHOTOMETRIC else 13)

    @property
    def c(self):
        p = _ffi.pba_problem()
        p.mode = self.mode
        p.n_poses, p.n_calib, p.n_landmarks, p.n_obs = self.n_poses, self.n_calib, self.n_landmarks, self.n_obs
        p.poses = _ffi.ptr(self.poses, C.c_double)
        p.pose_fixed = _ffi.ptr(self.pose_fixed, C.c_uint8)
        p.pose_calib = _ffi.ptr(self.pose_calib, C.c_int32)
        p.calib_model = _ffi.ptr(self.calib_model, C.c_int32)
        p.intrinsics = _ffi.ptr(self.intrinsics, C.c_double)
        p.inv_depth = _ffi.ptr(self.inv_depth, C.c_double)
        p.lm_host = _ffi.ptr(self.lm_host, C.c_int32)
        p.lm_host_uv = _ffi.ptr(self.lm_host_uv, C.c_double)
        p.lm_obs_ptr = _ffi.ptr(self.lm_obs_ptr, C.c_int64)
        p.obs_target = _ffi.ptr(self.obs_target, C.c_int32)
        p.obs_uv = _ffi.ptr(self.obs_uv, C.c_double)
        if self.images is not None:
            n, h, w = self.images.shape
            p.images = _ffi.ptr(self.images, C.c_uint8)
            p.image_stride = h * w
            p.width, p.height, p.pitch = w, h, w
        p.affine = _ffi.ptr(self.affine, C.c_double)
        self._c = p
        return p

    def copy(self):
        return Problem(self.mode, self.poses.copy(), self.pose_fixed, self.pose_calib, self.calib_model,
                       self.intrinsics, self.inv_depth.copy(), self.lm_host, self.lm_host_uv, self.lm_obs_ptr,
                       self.obs_target, self.obs_uv, self.images,
                       None if self.affine is None else self.affine.copy())

    def subset_landmarks(self, lo, hi):
        """Problem restricted to landmarks [lo, hi) (all poses kept)."""
        o0, o1 = int(self.lm_obs_ptr[lo]), int(self.lm_obs_ptr[hi])
        return Problem(self.mode, self.poses.copy(), self.pose_fixed, self.pose_calib, self.calib_model,
                       self.intrinsics, self.inv_depth[lo:hi].copy(), self.lm_host[lo:hi], self.lm_host_uv[lo:hi],
                       self.lm_obs_ptr[lo:hi + 1] - o0, self.obs_target[o0:o1],
                       None if self.obs_uv is None else self.obs_uv[o0:o1], self.images,
                       None if self.affine is None else self.affine.copy())


    def prefix_keyframes(self, n_kf):
        """The first n_kf keyframes and every landmark whose whole track lies inside them (landmarks must be
        ordered by host, as the synthetic scenes are); state copied."""
        if n_kf >= self.n_poses:
            return self.copy()
        last_target = np.maximum.reduceat(self.obs_target, self.lm_obs_ptr[:-1].clip(max=max(self.n_obs - 1, 0)))
        inside = (last_target < n_kf) & (self.lm_host < n_kf)
        L = int(np.argmin(inside)) if not inside.all() else self.n_landmarks
        o1 = int(self.lm_obs_ptr[L])
        return Problem(self.mode, self.poses[:n_kf].copy(), self.pose_fixed[:n_kf], self.pose_calib[:n_kf],
                       self.calib_model, self.intrinsics, self.inv_depth[:L].copy(), self.lm_host[:L],
                       self.lm_host_uv[:L], self.lm_obs_ptr[:L + 1], self.obs_target[:o1],
                       None if self.obs_uv is None else self.obs_uv[:o1],
                       None if self.images is None else self.images[:n_kf],
                       None if self.affine is None else self.affine[:n_kf].copy())

    def select_landmarks(self, lm_index):
        """Problem made of the given landmarks only (all poses kept, state shared by value) plus the
        caller-order observation indices of their blocks in this problem."""
        lm_index = np.asarray(lm_index, np.int64)
        lo, hi = self.lm_obs_ptr[lm_index], self.lm_obs_ptr[lm_index + 1]
        cnt = hi - lo
        ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        obs = np.repeat(lo - ptr[:-1], cnt) + np.arange(int(ptr[-1]), dtype=np.int64)
        sub = Problem(self.mode, self.poses.copy(), self.pose_fixed, self.pose_calib, self.calib_model,
                      self.intrinsics, self.inv_depth[lm_index].copy(), self.lm_host[lm_index],
                      self.lm_host_uv[lm_index], ptr, self.obs_target[obs],
                      None if self.obs_uv is None else self.obs_uv[obs], self.images,
                      None if self.affine is None else self.affine.copy())
        return sub, obs


def partition_landmarks(lm_obs_ptr, world_size):
    """Contiguous landmark ranges balanced by observation count.

    Returns world_size+1 boundaries.  The same rule is implemented in
    csrc/pba_host.cu:partition_landmarks (tests check they agree)."""
    lm_obs_ptr = np.asarray(lm_obs_ptr, np.int64)
    n_lm = lm_obs_ptr.shape[0] - 1
    total = int(lm_obs_ptr[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = (total * r) // world_size
        b = int(np.searchsorted(lm_obs_ptr, target, side="left"))
        b = min(max(b, bounds[-1]), n_lm)
        bounds.append(b)
    bounds.append(n_lm)
    return bounds
