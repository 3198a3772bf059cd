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


def _quat_mul(a, b):
    """Hamilton product of [x y z w] quaternions (rows)."""
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


def _quat_rot(q, v):
    """Rotate vectors v by unit quaternions q = [x y z w]."""
    u, w = q[..., :3], q[..., 3:4]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def _se3_plus(T, d):
    """T * exp(d), d = (upsilon, omega), rows [qx qy qz qw tx ty tz] (local_parameterization_se3.hpp:44-51)."""
    ups, om = d[:, :3], d[:, 3:]
    th = np.linalg.norm(om, axis=1, keepdims=True)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    dq = np.concatenate([np.where(small, 0.5, np.sin(0.5 * ths) / ths) * om, np.cos(0.5 * th)], 1)
    a = np.where(small, 0.5, (1.0 - np.cos(ths)) / ths ** 2)
    b = np.where(small, 1.0 / 6.0, (ths - np.sin(ths)) / ths ** 3)
    wu = np.cross(om, ups)
    vu = ups + a * wu + b * np.cross(om, wu)
    out = np.empty_like(T)
    q = _quat_mul(T[:, :4], dq)
    out[:, :4] = q / np.linalg.norm(q, axis=1, keepdims=True)
    out[:, 4:] = T[:, 4:] + _quat_rot(T[:, :4], vu)
    return out


def make_grid_scene(rows, cols, n_pts, spacing=1.5, track_len=9, seed=7, pose_sigma=0.01, rho_sigma=0.05,
                    pixel_sigma=0.3, width=752, height=480):
    """Geometric pinhole scene whose reduced camera system is NOT banded under any camera order: a lawn-mower
    (aerial-survey) flight, rows x cols keyframes on a planar grid `spacing` metres apart, all looking at the wall
    of SURVEY.md 8(d).  A landmark is hosted by a random keyframe and observed by up to track_len - 1 of the other
    keyframes that see it, so a keyframe is covisible with its neighbours along BOTH grid directions (loop closures
    between every pair of adjacent flight lines): the covisibility graph is a 2-D mesh whose bandwidth after
    reverse Cuthill-McKee stays ~ min(rows, cols) x the view radius.  Keyframe labels follow the flight (boustrophedon).
    Returns (Problem with the perturbed initial state, ground-truth dict), like make_scene."""
    rng = np.random.default_rng(seed)
    n_kf = rows * cols
    intr = np.array([370.34, 370.34, 375.5, 239.5, 0, 0, 0, 0.0])
    fx, fy, cx, cy = intr[:4]
    r, c = np.divmod(np.arange(n_kf), cols)
    c = np.where(r % 2 == 1, cols - 1 - c, c)  # boustrophedon
    gt = np.zeros((n_kf, 7))
    gt[:, 3] = 1.0
    gt[:, 4], gt[:, 5] = spacing * c, spacing * r
    wob = 0.02 * np.stack([np.sin(0.07 * np.arange(n_kf)), np.cos(0.05 * np.arange(n_kf)), 0.5 * np.sin(0.03 * np.arange(n_kf))], 1)
    gt = _se3_plus(gt, np.concatenate([np.zeros((n_kf, 3)), wob], 1))
    host = np.sort(rng.integers(0, n_kf, n_pts)).astype(np.int32)
    uv = np.stack([rng.uniform(100, width - 100, n_pts), rng.uniform(60, height - 60, n_pts)], 1)
    m = np.stack([(uv[:, 0] - cx) / fx, (uv[:, 1] - cy) / fy, np.ones(n_pts)], 1)
    b = m / np.linalg.norm(m, axis=1, keepdims=True)
    d = _quat_rot(gt[host, :4], b)
    o = gt[host, 4:]
    s = (5.0 - o[:, 2]) / d[:, 2]
    for _ in range(8):  # ray / height-field intersection, as synth_scene.h ray_wall
        s = (5.0 + 0.5 * np.sin(0.8 * (o[:, 0] + s * d[:, 0])) * np.cos(0.6 * (o[:, 1] + s * d[:, 1])) - o[:, 2]) / d[:, 2]
    Xw = o + s[:, None] * d
    qc = gt[:, :4] * np.array([-1.0, -1.0, -1.0, 1.0])
    ptr = np.zeros(n_pts + 1, np.int64)
    tgt, ouv = [], []
    reach = int(np.ceil(5.0 * max(width / fx, height / fy) / spacing)) + 1
    for l in range(n_pts):
        hr, hc = divmod(int(host[l]), cols)
        if hr % 2 == 1:
            hc = cols - 1 - hc
        rr = np.arange(max(0, hr - reach), min(rows, hr + reach + 1))
        cc = np.arange(max(0, hc - reach), min(cols, hc + reach + 1))
        R, Cc = np.meshgrid(rr, cc, indexing="ij")
        cand = (R * cols + np.where(R % 2 == 1, cols - 1 - Cc, Cc)).ravel()
        cand = np.sort(cand[cand > host[l]])  # the reference's host is obs.begin() = the smallest id (map_utils.h:351-352)
        Xt = _quat_rot(qc[cand], Xw[l] - gt[cand, 4:])
        pu, pv = fx * Xt[:, 0] / Xt[:, 2] + cx, fy * Xt[:, 1] / Xt[:, 2] + cy
        vis = (Xt[:, 2] > 0.5) & (pu > 20) & (pu < width - 20) & (pv > 20) & (pv < height - 20)
        cand, pu, pv = cand[vis], pu[vis], pv[vis]
        if len(cand) > track_len - 1:
            keep = np.sort(rng.choice(len(cand), track_len - 1, replace=False))
            cand, pu, pv = cand[keep], pu[keep], pv[keep]
        tgt.append(cand.astype(np.int32))
        ouv.append(np.stack([pu, pv], 1))
        ptr[l + 1] = ptr[l] + len(cand)
    obs_target = np.concatenate(tgt)
    obs_uv = np.concatenate(ouv) + pixel_sigma * rng.standard_normal((len(obs_target), 2))
    fixed = np.zeros(n_kf, np.uint8)
    fixed[:2] = 1
    poses = gt.copy()
    poses[2:] = _se3_plus(gt[2:], pose_sigma * rng.standard_normal((n_kf - 2, 6)))
    rho_gt = 1.0 / s
    rho = rho_gt / (1.0 + rho_sigma * rng.standard_normal(n_pts))
    prob = Problem(_ffi.MODE_GEOMETRIC, poses, fixed, np.zeros(n_kf, np.int32), np.array([_ffi.CAM_PINHOLE], np.int32),
                   intr.reshape(1, 8), rho, host, uv, ptr, obs_target, obs_uv, None, None)
    return prob, {"poses": gt, "inv_depth": rho_gt}
