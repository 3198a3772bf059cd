"""Landmark::get_p (common_types.h:205-217),
compute_projections + set_outlier_flags (src/sfm.cpp:1928-2008) and the
keep/remove decision of remove_outlier_landmarks (src/sfm.cpp:2029-2100).

CPU part: the oracle restatement against golden vectors made from the
reference's own get_p / SE3::inverse / project (tests/golden/projections_*.npz)
and, when oracle/_ref is present, against the reference live.
GPU part: pba_landmark_positions / pba_compute_projections through the C ABI
against the oracle.  Tolerance: 1e-11 relative on the fp64 outputs (FMA
contraction only); flags and removal decisions bit-exact wherever the deciding
quantity is not within 1e-9 of its threshold.
"""
import ctypes as C

import numpy as np
import pytest

import golden_util as gu
import oracle_ffi as of
import pba_b200 as pb
from pba_b200 import _ffi

FILES = gu.projection_files()
RTOL = 1e-11


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    return float(np.abs(a[m] - b[m]).max() / max(np.abs(b[m]).max(), 1e-300)) if m.any() else 0.0


def load(path):
    g = np.load(path)
    prob = pb.Problem(int(g["mode"]), g["poses"].copy(), g["pose_fixed"], g["pose_calib"], g["calib_model"],
                      g["intrinsics"], g["inv_depth"].copy(), g["lm_host"], g["lm_host_uv"], g["lm_obs_ptr"],
                      g["obs_target"], g["obs_uv"])
    return prob, g


def decisive(out, thr, eps=1e-9):
    """Slots whose flag-deciding quantities are all clear of their thresholds."""
    e = out["reprojection_error"]
    d = np.linalg.norm(out["point_3d_c"], axis=1)
    z = out["point_3d_c"][:, 2]
    ok = np.abs(e - thr.reprojection_error_huge_pixel) > eps
    ok &= np.abs(e - thr.reprojection_error_normal_pixel) > eps
    ok &= np.abs(d - thr.camera_center_distance_meter) > eps
    ok &= np.abs(z - thr.z_coordinate_meter) > eps
    return ok


def check_same(a, b, thr, with_remove=True):
    assert rel(a["point_3d_c"], b["point_3d_c"]) < RTOL
    assert rel(a["point_reprojected"], b["point_reprojected"]) < RTOL
    assert rel(a["reprojection_error"], b["reprojection_error"]) < 1e-9
    ok = decisive(b, thr)
    assert ok.mean() > 0.99
    assert np.array_equal(np.asarray(a["outlier_flags"])[ok], np.asarray(b["outlier_flags"])[ok])
    if with_remove and ok.all():
        assert np.array_equal(a["landmark_remove"], b["landmark_remove"])
        assert bool(a["any_severe_outliers"]) == bool(b["any_severe_outliers"])


def python_removal(flags, slot_ptr):
    """remove_outlier_landmarks' decision written the reference's way (src/sfm.cpp:2039-2091)."""
    any_severe = bool(np.any(flags & ~np.uint32(pb.projections.OutlierReprojectionErrorNormal)))
    remove = np.zeros(len(slot_ptr) - 1, np.uint8)
    for l in range(len(slot_ptr) - 1):
        for f in flags[slot_ptr[l]:slot_ptr[l + 1]]:
            if f & pb.projections.OutlierReprojectionErrorHuge:
                remove[l] = 1
                break
            if (f & pb.projections.OutlierReprojectionErrorNormal) and not any_severe:
                remove[l] = 1
                break
            if f & pb.projections.OutlierCameraDistance:
                remove[l] = 1
                break
            if f & pb.projections.OutlierZCoordinate:
                remove[l] = 1
                break
    return remove, any_severe


# ------------------------------------------------------------------ CPU side --
def test_projection_fixtures_exist():
    assert len(FILES) == 4, "run tests/golden/make_golden.py --projections"


def test_flag_values_match_reference_enum():
    # include/visnav/common_types.h:277-285
    p = pb.projections
    assert (p.OutlierNone, p.OutlierReprojectionErrorHuge, p.OutlierReprojectionErrorNormal, p.OutlierCameraDistance,
            p.OutlierZCoordinate) == (0, 1, 2, 4, 8)
    t = pb.ProjectionThresholds()  # src/sfm.cpp:254-261
    assert (t.reprojection_error_huge_pixel, t.reprojection_error_normal_pixel, t.camera_center_distance_meter,
            t.z_coordinate_meter) == (40.0, 3.0, 0.1, 0.05)


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_projections_match_reference_golden(path):
    prob, g = load(path)
    thr = pb.ProjectionThresholds()
    out = of.compute_projections("oracle", prob, thr)
    gold = {k[4:]: g[k] for k in g.files if k.startswith("ref_")}
    check_same(out, gold, thr, with_remove=False)
    assert rel(of.landmark_positions("oracle", prob), g["ref_p_w"]) < RTOL
    # every flag class is exercised by the fixture
    for bit in range(4):
        assert np.any((gold["outlier_flags"] >> bit) & 1)
    remove, severe = python_removal(out["outlier_flags"], prob.lm_obs_ptr + np.arange(prob.n_landmarks + 1))
    assert np.array_equal(remove, out["landmark_remove"]) and severe == out["any_severe_outliers"]


@pytest.mark.skipif(not of.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("model", ["pinhole", "ds", "kb4", "eucm"])
def test_oracle_projections_match_reference_live(model):
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 9, 400, model, seed_noise=11)
    thr = pb.ProjectionThresholds(1.5, 0.4, 2.5, 1.0)
    a = of.compute_projections("oracle", prob, thr)
    b = of.compute_projections("ref", prob, thr)
    check_same(a, b, thr, with_remove=False)
    assert rel(of.landmark_positions("oracle", prob), of.landmark_positions("ref", prob)) < RTOL


def test_removal_only_normal_outliers():
    """With no severe outlier anywhere, 'normal' reprojection outliers are removed (src/sfm.cpp:2067-2076);
    as soon as one severe outlier exists they are kept."""
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 6, 120, "pinhole", seed_noise=5)
    thr = pb.ProjectionThresholds(1e9, 0.3, 0.0, -1e9)
    out = of.compute_projections("oracle", prob, thr)
    assert not out["any_severe_outliers"] and out["landmark_remove"].any()
    sp = prob.lm_obs_ptr + np.arange(prob.n_landmarks + 1)
    normal = np.array([np.any(out["outlier_flags"][sp[l]:sp[l + 1]] & 2) for l in range(prob.n_landmarks)])
    assert np.array_equal(out["landmark_remove"].astype(bool), normal)
    thr2 = pb.ProjectionThresholds(1e9, 0.3, 0.0, 1e9)  # every slot gets the z flag
    out2 = of.compute_projections("oracle", prob, thr2)
    assert out2["any_severe_outliers"] and out2["landmark_remove"].all()


# ------------------------------------------------------------------ GPU side --
def gpu_projections(prob, thr):
    p = pb.compute_projections(prob, thr)
    return dict(point_reprojected=p.point_reprojected, point_3d_c=p.point_3d_c,
                reprojection_error=p.reprojection_error, outlier_flags=p.outlier_flags,
                landmark_remove=p.landmark_remove, any_severe_outliers=p.any_severe_outliers), p


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_projections_match_reference_golden(path):
    prob, g = load(path)
    thr = pb.ProjectionThresholds()
    out, _ = gpu_projections(prob, thr)
    gold = {k[4:]: g[k] for k in g.files if k.startswith("ref_")}
    check_same(out, gold, thr, with_remove=False)
    assert rel(pb.landmark_positions(prob), g["ref_p_w"]) < RTOL
    check_same(out, of.compute_projections("oracle", prob, thr), thr)


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["pinhole", "ds", "kb4", "eucm"])
@pytest.mark.parametrize("thr", [pb.ProjectionThresholds(), pb.ProjectionThresholds(1.5, 0.4, 2.5, 1.0),
                                 pb.ProjectionThresholds(1e9, 0.3, 0.0, -1e9)], ids=["default", "tight", "normal-only"])
def test_cuda_projections_match_oracle(model, thr):
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 12, 3000, model, seed_noise=3)
    out, p = gpu_projections(prob, thr)
    ref = of.compute_projections("oracle", prob, thr)
    check_same(out, ref, thr)
    assert rel(pb.landmark_positions(prob), of.landmark_positions("oracle", prob)) < RTOL
    l = int(np.argmax(ref["landmark_remove"])) if ref["landmark_remove"].any() else 0
    sp = p.slot_ptr
    assert p.is_landmark_outlier(l) == bool(np.any(ref["outlier_flags"][sp[l]:sp[l + 1]]))


@pytest.mark.gpu
def test_cuda_projections_mixed_models_and_ragged():
    """Two calibrations with different models (each observation projected with the OBSERVING camera's model,
    src/sfm.cpp:1974), landmarks with no observation besides the host, optional outputs left NULL."""
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 10, 800, "ds", seed_noise=9)
    intr = np.vstack([prob.intrinsics[0], [prob.intrinsics[0, 0] * 1.02, prob.intrinsics[0, 1] * 0.98,
                                           prob.intrinsics[0, 2], prob.intrinsics[0, 3], 0.55, 1.05, 0, 0]])
    pose_calib = (np.arange(prob.n_poses) % 2).astype(np.int32)
    # drop all non-host observations of every 7th landmark
    keep = np.ones(prob.n_obs, bool)
    for l in range(0, prob.n_landmarks, 7):
        keep[prob.lm_obs_ptr[l]:prob.lm_obs_ptr[l + 1]] = False
    cnt = np.diff(prob.lm_obs_ptr).copy()
    cnt[0::7] = 0
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    mixed = pb.Problem(prob.mode, prob.poses, prob.pose_fixed, pose_calib, [pb.CAM_DS, pb.CAM_EUCM], intr,
                       prob.inv_depth, prob.lm_host, prob.lm_host_uv, ptr, prob.obs_target[keep], prob.obs_uv[keep])
    thr = pb.ProjectionThresholds(30.0, 2.0, 0.1, 0.05)
    out, _ = gpu_projections(mixed, thr)
    check_same(out, of.compute_projections("oracle", mixed, thr), thr)
    # host slot of a landmark reprojects onto its own pixel (get_p round trip)
    sp = mixed.lm_obs_ptr + np.arange(mixed.n_landmarks + 1)
    assert np.abs(out["point_reprojected"][sp[:-1]] - mixed.lm_host_uv).max() < 1e-8
    # flags only
    lib = _ffi.load_lib()
    flags = np.zeros(mixed.n_obs + mixed.n_landmarks, np.uint32)
    t = thr.to_c()
    pc = mixed.c
    _ffi.check(lib.pba_compute_projections(C.byref(pc), C.byref(t), 0, None, None, None, _ffi.ptr(flags, C.c_uint32),
                                           None, None))
    assert np.array_equal(flags, out["outlier_flags"])


def test_projections_argument_errors_and_no_cpu_fallback():
    lib = _ffi.load_lib()
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 6, 50, "pinhole", seed_noise=1)
    t = pb.ProjectionThresholds().to_c()
    bad = prob.copy()
    bad.obs_target[0] = prob.n_poses  # out of range
    bad._c = None
    pc = bad.c
    assert lib.pba_compute_projections(C.byref(pc), C.byref(t), 0, None, None, None, None, None, None) == \
        1
    photo, _ = pb.make_scene(pb.MODE_PHOTOMETRIC, 6, 50, "pinhole", seed_noise=1)  # no corner positions
    pc = photo.c
    assert lib.pba_compute_projections(C.byref(pc), C.byref(t), 0, None, None, None, None, None, None) == \
        1
    pc = prob.c
    assert lib.pba_landmark_positions(C.byref(pc), 0, None) == 1
    if lib.pba_device_count() == 0:  # no CPU path: valid arguments without a GPU fail loudly
        out = np.zeros((prob.n_landmarks, 3))
        assert lib.pba_landmark_positions(C.byref(pc), 0, _ffi.ptr(out, C.c_double)) == 2  # PBA_ERR_NO_DEVICE
        with pytest.raises(RuntimeError):
            pb.compute_projections(prob)


@pytest.mark.gpu
def test_projections_after_solve_flag_nothing():
    """The SfM loop: optimize() then compute_projections(): a converged noise-free scene has no outliers."""
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 10, 1500, "kb4", seed_noise=21, pixel_sigma=0.0)
    pb.bundle_adjustment(prob, pb.BundleAdjustmentOptions(verbosity_level=0, max_num_iterations=30))
    p = pb.compute_projections(prob)
    assert p.reprojection_error.max() < 1e-3
    assert not p.outlier_flags.any() and not p.landmark_remove.any() and not p.any_severe_outliers


def _dropin_lib():
    import os
    base = os.path.dirname(of.REF_SO)
    path = os.path.join(base, "libpba_dropin_v4.so" if of.REF_SO.endswith("_v4.so") else "libpba_dropin.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    d = _ffi.c_double_p
    lib.pba_dropin_compute_projections.argtypes = [
        C.POINTER(_ffi.pba_problem), C.POINTER(_ffi.pba_projection_thresholds), C.c_int, d, d, d, d, _ffi.c_u32_p,
        _ffi.c_i64_p, _ffi.c_u8_p]
    return lib


def _dropin_projections(lib, prob, thr, use_b200):
    ns, nl = prob.n_obs + prob.n_landmarks, prob.n_landmarks
    out = dict(point_measured=np.zeros((ns, 2)), point_reprojected=np.zeros((ns, 2)), point_3d_c=np.zeros((ns, 3)),
               reprojection_error=np.zeros(ns), outlier_flags=np.zeros(ns, np.uint32),
               n_image_obs=np.zeros(prob.n_poses, np.int64), landmark_remove=np.zeros(nl, np.uint8))
    t = thr.to_c()
    pc = prob.c
    rc = lib.pba_dropin_compute_projections(
        C.byref(pc), C.byref(t), int(use_b200), _ffi.ptr(out["point_measured"], C.c_double),
        _ffi.ptr(out["point_reprojected"], C.c_double), _ffi.ptr(out["point_3d_c"], C.c_double),
        _ffi.ptr(out["reprojection_error"], C.c_double), _ffi.ptr(out["outlier_flags"], C.c_uint32),
        _ffi.ptr(out["n_image_obs"], C.c_int64), _ffi.ptr(out["landmark_remove"], C.c_uint8))
    assert rc == 0, rc
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["pinhole", "ds", "kb4", "eucm"])
def test_reference_containers_drive_both_projection_paths(model):
    """The reference's own Corners/Cameras/Landmarks/Calibration, built once, fill the reference's
    ImageProjections/TrackProjections (a) the way src/sfm.cpp:1960-1984 does with the reference's own
    get_p/inverse/project and (b) through include/visnav_b200/bundle_adjustment.h ->
    pba_compute_projections -> CUDA.  Same containers out."""
    lib = _dropin_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libpba_dropin.so not built on this box")
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 12, 900, model, seed_noise=17)
    for l in range(prob.n_landmarks):  # slot order = std::map<FrameCamId> order
        assert np.all(np.diff(prob.obs_target[prob.lm_obs_ptr[l]:prob.lm_obs_ptr[l + 1]]) > 0)
    thr = pb.ProjectionThresholds(6.0, 0.8, 2.5, 1.0)
    a = _dropin_projections(lib, prob, thr, 0)
    b = _dropin_projections(lib, prob, thr, 1)
    assert np.array_equal(a["point_measured"], b["point_measured"])
    assert np.array_equal(a["n_image_obs"], b["n_image_obs"]) and a["n_image_obs"].sum() == prob.n_obs + prob.n_landmarks
    check_same(b, a, thr, with_remove=False)
    ref = of.compute_projections("oracle", prob, thr)
    if decisive(ref, thr).all():
        assert np.array_equal(b["landmark_remove"], ref["landmark_remove"])
    assert a["outlier_flags"].any()
