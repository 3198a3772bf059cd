"""add_new_landmarks_between_cams (include/visnav/map_utils.h:121-195) — unit bearings of the
shared tracks, opengv's linear triangulation in camera 0's frame, initial inverse distance 1 / |p|.

Golden vectors (tests/golden/triangulate_*.npz, made by tests/golden/make_golden_triangulate.py) come from the
reference's OWN function + the vendored opengv.  CPU: the oracle restatement against them (and against oracle/_ref
live when present).  GPU: pba_triangulate_inverse_depth through the C ABI against both.
"""
import glob
import os

import numpy as np
import pytest

import oracle_ffi as of
import pba_b200 as pb

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "triangulate_*.npz")))
RTOL = 1e-9  # the null vector of a noisy 4x4 DLT matrix is determined to ~ eps * cond; 1e-9 relative is the bar


def rel_rows(a, b):
    return float((np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)).max())


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    m = int(g["model"])
    p, rho = of.triangulate("oracle", m, g["intr0"], m, g["intr1"], g["T_w_c0"], g["T_w_c1"], g["uv0"], g["uv1"])
    assert np.abs(rho - g["ref_inv_depth"]).max() <= 1e-11 * g["ref_inv_depth"].max()
    assert rel_rows(p, g["ref_p_c0"]) <= 1e-11
    if of.have_ref():
        p_r, rho_r = of.triangulate("ref", m, g["intr0"], m, g["intr1"], g["T_w_c0"], g["T_w_c1"], g["uv0"], g["uv1"])
        assert np.array_equal(rho_r, g["ref_inv_depth"]) or np.abs(rho_r - g["ref_inv_depth"]).max() <= 1e-13


def test_golden_files_exist():
    assert len(GOLDEN) == 4


def test_inverse_distance_is_measured_in_camera_zero():
    """map_utils.h:190 (the author's own TODO): rho = 1 / |p| with p in camera 0's frame."""
    g = np.load(GOLDEN[0])
    assert np.allclose(g["ref_inv_depth"], 1.0 / np.linalg.norm(g["ref_p_c0"], axis=1), rtol=1e-14)
    # and the triangulated points are near the true ones (0.3 px noise, 11 cm baseline)
    near = np.linalg.norm(g["true_p_c0"], axis=1) < 3.0
    assert np.abs(g["ref_p_c0"][near] - g["true_p_c0"][near]).max() < 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_triangulation_matches_reference_and_oracle(path):
    g = np.load(path)
    m = int(g["model"])
    p, rho = pb.triangulate_inverse_depth(m, g["intr0"], m, g["intr1"], g["T_w_c0"], g["T_w_c1"], g["uv0"], g["uv1"])
    assert np.abs(rho - g["ref_inv_depth"]).max() <= RTOL * g["ref_inv_depth"].max()
    assert rel_rows(p, g["ref_p_c0"]) <= RTOL
    p_o, rho_o = of.triangulate("oracle", m, g["intr0"], m, g["intr1"], g["T_w_c0"], g["T_w_c1"], g["uv0"], g["uv1"])
    assert np.abs(rho - rho_o).max() <= RTOL * rho_o.max()


@pytest.mark.gpu
def test_cuda_triangulation_mixed_models_and_many_points():
    """Different camera models on the two sides (cam 0 double sphere, cam 1 KB4), 200k tracks: CUDA == oracle."""
    import ctypes as C
    from pba_b200 import _ffi
    g0, g1 = np.load(GOLDEN[0]), np.load(GOLDEN[-1])
    rng = np.random.default_rng(1)
    n = 200000
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(1.0, 12.0, n)]
    T0 = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    T1 = np.array([0, 0, 0, 1, 0.2, 0.01, -0.02], np.float64)
    intr0 = np.array([0.5 * 805, 0.5 * 800, 505, 509, 0.5 * -0.150694, 0.5 * 1.48785, 0, 0])
    intr1 = np.array([379.045, 379.008, 505.512, 509.969, 0.00693023, -0.0013828, -0.000272596, -0.000452646])
    uv0, uv1 = np.zeros((n, 2)), np.zeros((n, 2))
    of.oracle().pba_oracle_project(pb.CAM_DS, _ffi.ptr(intr0, C.c_double), n, _ffi.ptr(X, C.c_double), _ffi.ptr(uv0, C.c_double), None)
    X1 = np.ascontiguousarray(X - T1[4:])
    of.oracle().pba_oracle_project(pb.CAM_KB4, _ffi.ptr(intr1, C.c_double), n, _ffi.ptr(X1, C.c_double), _ffi.ptr(uv1, C.c_double), None)
    uv0 += rng.normal(0, 0.2, uv0.shape); uv1 += rng.normal(0, 0.2, uv1.shape)
    p, rho = pb.triangulate_inverse_depth(pb.CAM_DS, intr0, pb.CAM_KB4, intr1, T0, T1, uv0, uv1)
    p_o, rho_o = of.triangulate("oracle", pb.CAM_DS, intr0, pb.CAM_KB4, intr1, T0, T1, uv0, uv1)
    assert np.all(np.abs(rho - rho_o) <= 1e-8 * np.abs(rho_o))
    assert np.median(np.abs(p - X).max(axis=1)) < 0.5  # 0.2 px noise, 20 cm baseline, points up to 12 m away
    del g0, g1


@pytest.mark.gpu
def test_reference_containers_drive_both_add_new_landmarks():
    """Drop-in proof: the reference's own containers through the reference's add_new_landmarks_between_cams and
    through visnav_b200's (include/visnav_b200/bundle_adjustment.h): same landmarks added, same inverse distances,
    same observation sets; existing landmarks and unshared tracks untouched."""
    import ctypes as C
    from pba_b200 import _ffi
    base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    path = os.path.join(base, "libpba_dropin_v4.so" if of.REF_SO.endswith("_v4.so") else "libpba_dropin.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libpba_dropin.so not built on this box")
    _ffi.load_lib()
    lib = C.CDLL(path)
    d = _ffi.c_double_p
    lib.pba_dropin_add_new_landmarks.argtypes = [C.c_int, d, C.c_int, d, d, d, C.c_int64, d, d, C.c_int, d, _ffi.c_i32_p,
                                                 _ffi.c_i32_p]
    g = np.load(GOLDEN[1])
    m = int(g["model"])
    n = g["uv0"].shape[0]
    out = []
    for use_b200 in (0, 1):
        rho, nobs, added = np.zeros(n), np.zeros(n, np.int32), C.c_int32()
        uv0, uv1 = np.ascontiguousarray(g["uv0"]), np.ascontiguousarray(g["uv1"])
        rc = lib.pba_dropin_add_new_landmarks(m, _ffi.ptr(g["intr0"], C.c_double), m, _ffi.ptr(g["intr1"], C.c_double),
                                              _ffi.ptr(g["T_w_c0"], C.c_double), _ffi.ptr(g["T_w_c1"], C.c_double), n,
                                              _ffi.ptr(uv0, C.c_double), _ffi.ptr(uv1, C.c_double), use_b200,
                                              _ffi.ptr(rho, C.c_double), _ffi.ptr(nobs, C.c_int32), C.byref(added))
        assert rc == 0
        out.append((rho, nobs, added.value))
    (rho_r, nobs_r, add_r), (rho_g, nobs_g, add_g) = out
    idx = np.arange(n)
    expect_new = (idx % 3 != 0) & (idx % 5 != 4)
    assert add_r == add_g == int(expect_new.sum())
    assert np.array_equal(nobs_r, nobs_g)
    assert np.all(rho_g[idx % 3 == 0] == -1.0) and np.all(rho_g[(idx % 3 != 0) & (idx % 5 == 4)] == -2.0)
    assert np.abs(rho_g[expect_new] - rho_r[expect_new]).max() <= RTOL * rho_r[expect_new].max()
    assert np.all(nobs_g[expect_new] == 2)  # the camera that is not in the map is not an observation
