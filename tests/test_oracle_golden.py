"""The CPU oracle (oracle/pba_oracle.cpp) against the golden vectors produced by
the REAL reference (tests/golden/*.npz, made by tests/golden/make_golden.py from
oracle/_ref = unmodified visnav headers + vendored Ceres 2.0.0), and — when the
built reference is present — against the reference live."""
import numpy as np
import pytest

import golden_util as gu
import oracle_ffi as of
import pba_b200 as pb

FILES = gu.golden_files()


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_fixtures_exist():
    assert len(FILES) == 8, "run tests/golden/make_golden.py"


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_residuals_jacobians_match_reference_golden(path):
    prob, g = gu.load(path)
    hub = float(g["huber"])
    cost, r, J = of.evaluate("oracle", prob, True, hub, threads=2)
    assert abs(cost - float(g["ref_cost"])) <= 1e-12 * float(g["ref_cost"])
    assert rel(r, g["ref_residuals"]) < 1e-11
    assert rel(J, g["ref_jacobians"]) < 1e-11
    cost, r, J = of.evaluate("oracle", prob, False, hub, threads=2)
    assert rel(r, g["ref_residuals_nohuber"]) < 1e-11
    assert rel(J, g["ref_jacobians_nohuber"]) < 1e-11


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_lm_matches_reference_golden(path):
    prob, g = gu.load(path)
    s = of.solve("oracle", prob, of.default_options(huber_parameter=float(g["huber"])), threads=2)
    assert s.termination_type == int(g["sol_termination"])
    assert s.num_iterations == len(g["sol_iter_cost"])
    assert abs(s.final_cost - float(g["sol_final_cost"])) <= 1e-6 * float(g["sol_final_cost"])
    its = s.iterations
    np.testing.assert_allclose([i["cost"] for i in its], g["sol_iter_cost"], rtol=1e-6)
    np.testing.assert_allclose([i["trust_region_radius"] for i in its], g["sol_iter_radius"], rtol=1e-6)
    assert [i["step_is_successful"] for i in its] == list(g["sol_iter_success"])
    np.testing.assert_allclose([i["gradient_max_norm"] for i in its], g["sol_iter_gradient_max_norm"], rtol=1e-4, atol=1e-9)
    assert np.abs(prob.poses - g["sol_poses"]).max() < 1e-5
    assert np.abs(prob.inv_depth - g["sol_inv_depth"]).max() < 1e-5
    if "sol_affine" in g:
        assert np.abs(prob.affine - g["sol_affine"]).max() < 1e-4
    if "entry_poses" in g:  # the unmodified visnav::bundle_adjustment() entry point
        assert np.abs(prob.poses - g["entry_poses"]).max() < 1e-5
        assert np.abs(prob.inv_depth - g["entry_inv_depth"]).max() < 1e-5


@pytest.mark.skipif(not of.have_ref(), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
@pytest.mark.parametrize("model", ["pinhole", "ds", "kb4", "eucm"])
def test_oracle_matches_reference_live(mode, model):
    prob, _ = pb.make_scene(mode, 7, 120, model)
    hub = 9.0 if mode else 1.0
    c1, r1, J1 = of.evaluate("ref", prob, True, hub, threads=2)
    c2, r2, J2 = of.evaluate("oracle", prob, True, hub, threads=2)
    assert abs(c1 - c2) <= 1e-12 * c1 and rel(r2, r1) < 1e-11 and rel(J2, J1) < 1e-11
    pa, pb_ = prob.copy(), prob.copy()
    sa = of.solve("ref", pa, of.default_options(huber_parameter=hub), threads=2)
    sb = of.solve("oracle", pb_, of.default_options(huber_parameter=hub), threads=2)
    assert sa.num_iterations == sb.num_iterations and sa.termination_type == sb.termination_type
    assert abs(sa.final_cost - sb.final_cost) <= 1e-6 * sa.final_cost
    assert np.abs(pa.poses - pb_.poses).max() < 1e-5


@pytest.mark.skipif(not of.have_ref(), reason="oracle/_ref not built on this box")
def test_reference_camera_models_and_se3_pin_the_oracle():
    import ctypes as C
    from pba_b200 import _ffi
    rng = np.random.default_rng(0)
    n = 500
    xyz = np.c_[rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(0.5, 5, n)]
    # getTestProjections() intrinsics (camera_models.h:60-66,134-140,211-218,300-307)
    intrs = {0: [402.5, 400, 505, 509, 0, 0, 0, 0], 3: [250, 250, 319.5, 239.5, 0.51231234, 0.9, 0, 0],
             1: [402.5, 400, 505, 509, -0.075347, 0.743925, 0, 0],
             2: [379.045, 379.008, 505.512, 509.969, 0.00693023, -0.0013828, -0.000272596, -0.000452646]}
    for model, intr in intrs.items():
        intr = np.array(intr, np.float64)
        a, b = np.zeros((n, 2)), np.zeros((n, 2))
        Ja, Jb = np.zeros((n, 6)), np.zeros((n, 6))
        of.ref().pba_ref_project(model, _ffi.ptr(intr, C.c_double), n, _ffi.ptr(xyz, C.c_double), _ffi.ptr(a, C.c_double))
        of.ref().pba_ref_project_jacobian(model, _ffi.ptr(intr, C.c_double), n, _ffi.ptr(xyz, C.c_double), _ffi.ptr(Ja, C.c_double))
        of.oracle().pba_oracle_project(model, _ffi.ptr(intr, C.c_double), n, _ffi.ptr(xyz, C.c_double),
                                       _ffi.ptr(b, C.c_double), _ffi.ptr(Jb, C.c_double))
        assert rel(b, a) < 1e-14 and rel(Jb, Ja) < 1e-13
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    T = np.c_[q, rng.normal(size=(n, 3))]
    d = rng.normal(scale=0.2, size=(n, 6)); d[0] = 0; d[1, 3:] = 1e-13
    oa, ob = np.zeros((n, 7)), np.zeros((n, 7))
    of.ref().pba_ref_se3_plus(n, _ffi.ptr(T, C.c_double), _ffi.ptr(d, C.c_double), _ffi.ptr(oa, C.c_double))
    of.oracle().pba_oracle_se3_plus(n, _ffi.ptr(T, C.c_double), _ffi.ptr(d, C.c_double), _ffi.ptr(ob, C.c_double))
    assert np.abs(oa - ob).max() < 1e-14
    Ja, Jb = np.zeros(42), np.zeros(42)
    of.ref().pba_ref_se3_plus_jacobian(_ffi.ptr(T[3], C.c_double), _ffi.ptr(Ja, C.c_double))
    of.oracle().pba_oracle_se3_plus_jacobian(_ffi.ptr(T[3], C.c_double), _ffi.ptr(Jb, C.c_double))
    assert np.abs(Ja - Jb).max() < 1e-15
