"""Deterministic synthetic BA scenes; wraps csrc/synth.cpp."""
import ctypes as C

import numpy as np

from . import _ffi
from .problem import Problem


def make_scene(mode, n_kf, n_pts, model="pinhole", width=752, height=480, render=True, gpu_render=False,
               **overrides):
    """Returns (Problem with the perturbed initial state, ground-truth dict).

    gpu_render=True ray-casts the keyframe images on the current CUDA device
    (libpba_b200.so: pba_synth_render_gpu) — for the 2,000-keyframe benchmark."""
    s = _ffi.load_synth()
    model_id = _ffi.CAM_NAMES[model] if isinstance(model, str) else int(model)
    prm = _ffi.pba_synth_params()
    s.pba_synth_default_params(C.byref(prm), int(mode), int(n_kf), int(n_pts), model_id)
    prm.width, prm.height = width, height
    for k, v in overrides.items():
        setattr(prm, k, v)
    n_obs = s.pba_synth_count_obs(C.byref(prm))
    photo = mode == _ffi.MODE_PHOTOMETRIC
    poses_gt = np.zeros((n_kf, 7))
    poses = np.zeros((n_kf, 7))
    fixed = np.zeros(n_kf, np.uint8)
    rho_gt = np.zeros(n_pts)
    rho = np.zeros(n_pts)
    lm_host = np.zeros(n_pts, np.int32)
    lm_host_uv = np.zeros((n_pts, 2))
    lm_obs_ptr = np.zeros(n_pts + 1, np.int64)
    obs_target = np.zeros(n_obs, np.int32)
    obs_uv = None if photo else np.zeros((n_obs, 2))
    affine = np.zeros((n_kf, 2)) if photo else None
    rc = s.pba_synth_generate(C.byref(prm), _ffi.ptr(poses_gt, C.c_double), _ffi.ptr(poses, C.c_double),
                              _ffi.ptr(fixed, C.c_uint8), _ffi.ptr(rho_gt, C.c_double), _ffi.ptr(rho, C.c_double),
                              _ffi.ptr(lm_host, C.c_int32), _ffi.ptr(lm_host_uv, C.c_double),
                              _ffi.ptr(lm_obs_ptr, C.c_int64), _ffi.ptr(obs_target, C.c_int32),
                              _ffi.ptr(obs_uv, C.c_double), _ffi.ptr(affine, C.c_double))
    if rc != 0:
        raise ValueError("pba_synth_generate failed (%d)" % rc)
    images = None
    if photo:
        images = np.zeros((n_kf, height, width), np.uint8)
        if render and gpu_render:
            _ffi.check(_ffi.load_lib().pba_synth_render_gpu(C.byref(prm), 0, n_kf, width,
                                                            _ffi.ptr(images, C.c_uint8)), "pba_synth_render_gpu")
        elif render:
            s.pba_synth_render(C.byref(prm), 0, n_kf, width, _ffi.ptr(images, C.c_uint8))
    intr = np.array(list(prm.intrinsics)).reshape(1, 8)
    prob = Problem(mode, poses, fixed, np.zeros(n_kf, np.int32), np.array([model_id], np.int32), intr, rho,
                   lm_host, lm_host_uv, lm_obs_ptr, obs_target, obs_uv, images, affine)
    gt = {"poses": poses_gt, "inv_depth": rho_gt, "params": prm}
    return prob, gt
