"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU
oracle (oracle/libpba_oracle.so, pinned to the real reference in
test_oracle_vs_reference.py) on the same seeded inputs.

Tolerances (BASELINE.json north_star): per-residual values and Jacobians
<= 1e-9 relative in fp64; after the LM run final cost <= 1e-6 relative and
poses / inverse depths <= 1e-5.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_ffi as of
import pba_b200 as pb
from pba_b200 import _ffi

pytestmark = pytest.mark.gpu

RTOL_RJ = 1e-9
RTOL_COST = 1e-6
TOL_STATE = 1e-5

MODELS = ["pinhole", "ds", "kb4", "eucm"]
# getTestProjections() intrinsics of the reference (camera_models.h:60-66,134-140,211-218,300-307)
TEST_INTR = {
    pb.CAM_PINHOLE: [0.5 * 805, 0.5 * 800, 505, 509, 0, 0, 0, 0],
    pb.CAM_EUCM: [0.5 * 500, 0.5 * 500, 319.5, 239.5, 0.51231234, 0.9, 0, 0],
    pb.CAM_DS: [0.5 * 805, 0.5 * 800, 505, 509, 0.5 * -0.150694, 0.5 * 1.48785, 0, 0],
    pb.CAM_KB4: [379.045, 379.008, 505.512, 509.969, 0.00693023, -0.0013828, -0.000272596, -0.000452646],
}


def rel(a, b):
    """max |a-b| / max |b|; NaNs (outside a model's domain — the reference has no
    domain checks) must sit at the same places in both."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    if not m.any():
        return 0.0
    return float(np.abs(a[m] - b[m]).max() / max(np.abs(b[m]).max(), 1e-300))


def scene(mode, model, n_kf=8, n_pts=300, **kw):
    return pb.make_scene(mode, n_kf, n_pts, model, **kw)


def huber_for(mode):
    return 9.0 if mode == pb.MODE_PHOTOMETRIC else 1.0


# ---------------------------------------------------------------- primitives --
@pytest.mark.parametrize("model", [pb.CAM_PINHOLE, pb.CAM_DS, pb.CAM_KB4, pb.CAM_EUCM])
def test_camera_models(model):
    lib = _ffi.load_lib()
    rng = np.random.default_rng(7)
    n = 4096
    xyz = np.c_[rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(0.5, 6, n)]
    xyz[0, :2] = 0.0  # r == 0 special case of KB4 (camera_models.h:333-337)
    intr = np.array(TEST_INTR[model], np.float64)
    uv = np.zeros((n, 2)); J = np.zeros((n, 6))
    _ffi.check(lib.pba_camera_project(model, _ffi.ptr(intr, C.c_double), n, _ffi.ptr(xyz, C.c_double),
                                      _ffi.ptr(uv, C.c_double), _ffi.ptr(J, C.c_double)))
    uv_o = np.zeros((n, 2)); J_o = np.zeros((n, 6))
    of.oracle().pba_oracle_project(model, _ffi.ptr(intr, C.c_double), n, _ffi.ptr(xyz, C.c_double),
                                   _ffi.ptr(uv_o, C.c_double), _ffi.ptr(J_o, C.c_double))
    assert rel(uv, uv_o) < 1e-12
    assert rel(J, J_o) < 1e-10
    # unproject on pixels, and project(unproject(p)) == p
    px = np.c_[rng.uniform(50, 700, n), rng.uniform(50, 430, n)]
    px[0] = intr[2:4]
    b = np.zeros((n, 3)); b_o = np.zeros((n, 3))
    _ffi.check(lib.pba_camera_unproject(model, _ffi.ptr(intr, C.c_double), n, _ffi.ptr(px, C.c_double),
                                        _ffi.ptr(b, C.c_double)))
    of.oracle().pba_oracle_unproject(model, _ffi.ptr(intr, C.c_double), n, _ffi.ptr(px, C.c_double),
                                     _ffi.ptr(b_o, C.c_double))
    assert rel(b, b_o) < 1e-12
    back = np.zeros((n, 2))
    _ffi.check(lib.pba_camera_project(model, _ffi.ptr(intr, C.c_double), n, _ffi.ptr(b, C.c_double),
                                      _ffi.ptr(back, C.c_double), None))
    ok = ~np.isnan(back).any(axis=1)
    assert ok.sum() > n // 2 and np.abs(back[ok] - px[ok]).max() < 1e-6


def test_se3_plus():
    lib = _ffi.load_lib()
    rng = np.random.default_rng(3)
    n = 2000
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    T = np.c_[q, rng.normal(size=(n, 3))]
    d = rng.normal(scale=0.3, size=(n, 6))
    d[0] = 0.0            # identity step
    d[1, 3:] = 1e-12      # small-angle branch (so3.hpp:596-603)
    d[2, 3:] = [np.pi, 0, 0]
    out = np.zeros((n, 7)); out_o = np.zeros((n, 7))
    _ffi.check(lib.pba_se3_plus(n, _ffi.ptr(T, C.c_double), _ffi.ptr(d, C.c_double), _ffi.ptr(out, C.c_double)))
    of.oracle().pba_oracle_se3_plus(n, _ffi.ptr(T, C.c_double), _ffi.ptr(d, C.c_double), _ffi.ptr(out_o, C.c_double))
    assert np.abs(out - out_o).max() < 1e-13
    assert np.abs(np.linalg.norm(out[:, :4], axis=1) - 1).max() < 1e-14


@pytest.mark.parametrize("n", [1, 8, 64, 100, 392, 1000])
def test_dense_cholesky_dmma(n):
    lib = _ffi.load_lib()
    rng = np.random.default_rng(n)
    M = rng.normal(size=(n, n))
    A = M @ M.T + n * np.eye(n)
    b = rng.normal(size=n)
    x = np.zeros(n)
    _ffi.check(lib.pba_cholesky_solve(n, _ffi.ptr(np.ascontiguousarray(A), C.c_double), _ffi.ptr(b, C.c_double),
                                      _ffi.ptr(x, C.c_double)))
    x_ref = np.linalg.solve(A, b)
    assert rel(x, x_ref) < 1e-11
    # not positive definite -> numerical failure status, no crash
    if n >= 8:
        A2 = A.copy(); A2[n // 2, n // 2] = -1.0
        st = lib.pba_cholesky_solve(n, _ffi.ptr(np.ascontiguousarray(A2), C.c_double), _ffi.ptr(b, C.c_double),
                                    _ffi.ptr(x, C.c_double))
        assert st == 5


# ------------------------------------------------- residuals and Jacobians --
@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("use_huber", [True, False])
def test_residual_jacobian_parity(mode, model, use_huber):
    prob, _ = scene(mode, model)
    hub = huber_for(mode)
    cost_o, r_o, J_o = of.evaluate("oracle", prob, use_huber, hub)
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(use_huber=use_huber, huber_parameter=hub))
    cost = eng.evaluate(True)
    r, J = eng.residuals(), eng.jacobians()
    assert abs(cost - cost_o) <= 1e-12 * abs(cost_o)
    assert rel(r, r_o) < RTOL_RJ
    assert rel(J, J_o) < RTOL_RJ
    # per-block check too, so a small block cannot hide behind a large one
    blk = np.abs(J_o).reshape(J_o.shape[0], -1).max(axis=1)
    err = np.abs(J - J_o).reshape(J_o.shape[0], -1).max(axis=1)
    assert (err <= RTOL_RJ * np.maximum(blk, 1e-12)).all()
    # cost-only kernel agrees with the Jacobian kernel
    assert abs(eng.evaluate(False) - cost) <= 1e-13 * abs(cost)
    eng.close()


def test_photometric_out_of_bounds_blocks_are_zero():
    prob, _ = scene(pb.MODE_PHOTOMETRIC, "pinhole", n_kf=6, n_pts=120)
    prob.lm_host_uv[:10] = [1.0, 1.0]            # host pattern leaves the image
    prob.inv_depth[10:20] *= 40.0                # projects far outside the target
    cost_o, r_o, J_o = of.evaluate("oracle", prob, True, 9.0)
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(huber_parameter=9.0))
    cost = eng.evaluate(True)
    r, J = eng.residuals(), eng.jacobians()
    zero_blocks = np.where(np.abs(r_o).max(axis=1) == 0)[0]
    assert len(zero_blocks) >= 10
    assert np.abs(r[zero_blocks]).max() == 0 and np.abs(J[zero_blocks]).max() == 0
    assert rel(r, r_o) < RTOL_RJ and rel(J, J_o) < RTOL_RJ
    assert abs(cost - cost_o) <= 1e-12 * abs(cost_o)
    eng.close()


# ------------------------------------------------------------ Schur / RCS ---
@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
@pytest.mark.parametrize("model", ["pinhole", "ds"])
def test_reduced_camera_system_parity(mode, model):
    prob, _ = scene(mode, model, n_kf=10, n_pts=500)
    hub = huber_for(mode)
    S_o, rhs_o, _ = of.build_rcs(prob, True, hub, radius=1e4)
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(huber_parameter=hub))
    eng.evaluate(True)
    eng.build_rcs(1e4)
    S, rhs = eng.rcs()
    assert S.shape == S_o.shape
    assert rel(S, S_o) < 1e-10
    assert rel(rhs, rhs_o) < 1e-10
    y_ref = np.linalg.solve(S_o, rhs_o)
    y_chol, _ = eng.solve_rcs(pb.SOLVER_CHOLESKY)
    assert rel(y_chol, y_ref) < 1e-8
    y_band, _ = eng.solve_rcs(pb.SOLVER_BAND)
    assert rel(y_band, y_ref) < 1e-8
    y_bcr, _ = eng.solve_rcs(pb.SOLVER_BCR)
    assert rel(y_bcr, y_ref) < 1e-8
    y_pcg, iters = eng.solve_rcs(pb.SOLVER_PCG)
    assert iters > 0
    assert rel(y_pcg, y_ref) < 1e-6
    eng.close()


@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
def test_reduced_camera_system_parity_wide_windows(mode):
    """Landmarks seen by 15-18 keyframes: host groups with up to 17 cameras (the generic Schur
    product kernel, not the row-paired one for <= 12) and a half-bandwidth too wide for the
    band / BCR solvers' shared-memory blocks (dense Cholesky / PCG take over)."""
    prob, _ = scene(mode, "pinhole", n_kf=40, n_pts=1200, min_len=15, max_len=18)
    assert np.diff(prob.lm_obs_ptr).max() >= 15
    hub = huber_for(mode)
    S_o, rhs_o, _ = of.build_rcs(prob, True, hub, radius=1e4)
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(huber_parameter=hub))
    eng.evaluate(True)
    eng.build_rcs(1e4)
    S, rhs = eng.rcs()
    assert rel(S, S_o) < 1e-10
    assert rel(rhs, rhs_o) < 1e-10
    y_ref = np.linalg.solve(S_o, rhs_o)
    y, _ = eng.solve_rcs(pb.SOLVER_AUTO)
    assert rel(y, y_ref) < 1e-8
    eng.close()


# ------------------------------------------------------------------ LM run --
@pytest.mark.parametrize("mode,model,n_kf,n_pts", [
    (pb.MODE_GEOMETRIC, "pinhole", 10, 400),
    (pb.MODE_GEOMETRIC, "ds", 12, 500),
    (pb.MODE_GEOMETRIC, "kb4", 10, 400),
    (pb.MODE_PHOTOMETRIC, "pinhole", 10, 400),
    (pb.MODE_PHOTOMETRIC, "ds", 10, 400),
    (pb.MODE_PHOTOMETRIC, "kb4", 10, 400),
])
def test_lm_matches_oracle(mode, model, n_kf, n_pts):
    prob, _ = scene(mode, model, n_kf=n_kf, n_pts=n_pts)
    hub = huber_for(mode)
    po = prob.copy()
    so = of.solve("oracle", po, of.default_options(huber_parameter=hub))
    pg = prob.copy()
    sg = pb.bundle_adjustment(pg, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub,
                                                             solver=pb.SOLVER_CHOLESKY))
    assert sg.termination_type == so.termination_type
    assert sg.num_iterations == so.num_iterations
    assert abs(sg.initial_cost - so.initial_cost) <= 1e-12 * so.initial_cost
    assert abs(sg.final_cost - so.final_cost) <= RTOL_COST * so.final_cost
    for a, b in zip(sg.iterations, so.iterations):
        assert a["step_is_successful"] == b["step_is_successful"]
        assert abs(a["cost"] - b["cost"]) <= RTOL_COST * b["cost"]
        assert abs(a["trust_region_radius"] - b["trust_region_radius"]) <= 1e-6 * b["trust_region_radius"]
    assert np.abs(pg.poses - po.poses).max() < TOL_STATE
    assert np.abs(pg.inv_depth - po.inv_depth).max() < TOL_STATE
    if mode == pb.MODE_PHOTOMETRIC:
        assert np.abs(pg.affine - po.affine).max() < 1e-4
    assert sg.gpu_kernel_launches > 0


@pytest.mark.parametrize("solver", [pb.SOLVER_PCG, pb.SOLVER_BAND, pb.SOLVER_BCR, pb.SOLVER_AUTO])
@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
def test_lm_other_solvers_reach_same_cost(solver, mode):
    prob, _ = scene(mode, "pinhole", n_kf=30, n_pts=1500)
    hub = huber_for(mode)
    po = prob.copy()
    so = of.solve("oracle", po, of.default_options(huber_parameter=hub))
    pg = prob.copy()
    sg = pb.bundle_adjustment(pg, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, solver=solver))
    # AUTO: a 30-keyframe chain is short enough for the sequential band factorisation (<= 64 free keyframes)
    assert sg.linear_solver == (pb.SOLVER_BAND if solver == pb.SOLVER_AUTO else solver)
    assert abs(sg.final_cost - so.final_cost) <= RTOL_COST * so.final_cost
    assert np.abs(pg.poses - po.poses).max() < TOL_STATE
    assert np.abs(pg.inv_depth - po.inv_depth).max() < TOL_STATE


def test_fixed_cameras_and_unobserved_blocks_untouched():
    prob, _ = scene(pb.MODE_GEOMETRIC, "pinhole", n_kf=9, n_pts=300)
    # a landmark with no observation contributes nothing and stays untouched (map_utils.h:355)
    ptr = prob.lm_obs_ptr.copy()
    n0 = int(ptr[1] - ptr[0])
    prob2 = pb.Problem(prob.mode, prob.poses, prob.pose_fixed, prob.pose_calib, prob.calib_model, prob.intrinsics,
                       prob.inv_depth, prob.lm_host, prob.lm_host_uv,
                       np.r_[0, 0, ptr[2:] - ptr[1]], prob.obs_target[ptr[1]:], prob.obs_uv[ptr[1]:])
    before = prob2.copy()
    s = pb.bundle_adjustment(prob2, pb.BundleAdjustmentOptions(verbosity_level=0))
    assert n0 > 0 and s.termination_type != pb.FAILURE
    fixed = prob2.pose_fixed.astype(bool)
    assert np.array_equal(prob2.poses[fixed], before.poses[fixed])
    assert prob2.inv_depth[0] == before.inv_depth[0]
    assert np.abs(prob2.poses[~fixed] - before.poses[~fixed]).max() > 0
    po = before.copy()
    so = of.solve("oracle", po, of.default_options())
    assert abs(s.final_cost - so.final_cost) <= RTOL_COST * so.final_cost


def test_error_behaviour():
    prob, _ = scene(pb.MODE_GEOMETRIC, "pinhole", n_kf=6, n_pts=50)
    lib = _ffi.load_lib()
    # optimize_intrinsics is rejected (the reference marks it broken, map_utils.h:339)
    o = pb.BundleAdjustmentOptions(verbosity_level=0, optimize_intrinsics=True).to_c()
    s = pb.Summary()
    pc = prob.c
    assert lib.pba_solve(C.byref(pc), C.byref(o), C.byref(s.c)) == 4
    # bad target index
    bad = prob.copy(); bad.obs_target = bad.obs_target.copy(); bad.obs_target[0] = 99
    pc = bad.c
    o = pb.BundleAdjustmentOptions(verbosity_level=0).to_c()
    assert lib.pba_solve(C.byref(pc), C.byref(o), C.byref(s.c)) == 1
    # empty problem: nothing to do, success
    empty = pb.Problem(pb.MODE_GEOMETRIC, prob.poses, prob.pose_fixed, prob.pose_calib, prob.calib_model,
                       prob.intrinsics, np.zeros(0), np.zeros(0, np.int32), np.zeros((0, 2)), np.zeros(1, np.int64),
                       np.zeros(0, np.int32), np.zeros((0, 2)))
    s2 = pb.bundle_adjustment(empty, pb.BundleAdjustmentOptions(verbosity_level=0))
    assert s2.initial_cost == 0.0 and s2.termination_type == pb.CONVERGENCE


# ------------------------------------------ size-independent properties -----
def test_properties_at_scale():
    """BASELINE config-2 shape (50 KF x 20k pts): properties that need no oracle."""
    prob, gt = pb.make_scene(pb.MODE_PHOTOMETRIC, 50, 20000, "pinhole")
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=9.0))
    c1 = eng.evaluate(True)
    c2 = eng.evaluate(True)
    assert c1 == c2                      # deterministic: no atomics on the path
    assert abs(eng.evaluate(False) - c1) <= 1e-13 * c1
    # the ground-truth state has (almost) zero photometric cost relative to the perturbed one
    eng.set_state(gt["poses"], gt["inv_depth"], np.zeros((50, 2)))
    c_gt = eng.evaluate(False)
    assert c_gt < 0.05 * c1
    eng.set_state(prob.poses, prob.inv_depth, prob.affine)
    eng.build_rcs(1e4)
    S, rhs = eng.rcs()
    assert np.array_equal(S, S.T)
    assert np.linalg.eigvalsh(S).min() > 0      # damped RCS is SPD
    s = eng.minimize()
    assert s.final_cost < 0.1 * s.initial_cost
    poses, rho, aff = eng.get_state()
    assert np.abs(poses - gt["poses"]).max() < np.abs(prob.poses - gt["poses"]).max()
    eng.close()


# ------------------------------------- golden vectors of the real reference --
import golden_util as gu  # noqa: E402


@pytest.mark.parametrize("path", gu.golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_matches_reference_golden_vectors(path):
    """CUDA path vs the committed outputs of the unmodified reference
    (visnav functor + vendored Ceres AutoDiff; tests/golden/make_golden.py)."""
    prob, g = gu.load(path)
    hub = float(g["huber"])
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub))
    cost = eng.evaluate(True)
    assert abs(cost - float(g["ref_cost"])) <= 1e-12 * float(g["ref_cost"])
    assert rel(eng.residuals(), g["ref_residuals"]) < RTOL_RJ
    assert rel(eng.jacobians(), g["ref_jacobians"]) < RTOL_RJ
    eng.close()
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(verbosity_level=0, use_huber=False))
    eng.evaluate(True)
    assert rel(eng.residuals(), g["ref_residuals_nohuber"]) < RTOL_RJ
    assert rel(eng.jacobians(), g["ref_jacobians_nohuber"]) < RTOL_RJ
    eng.close()
    s = pb.bundle_adjustment(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub,
                                                              solver=pb.SOLVER_CHOLESKY))
    assert s.termination_type == int(g["sol_termination"])
    assert s.num_iterations == len(g["sol_iter_cost"])
    assert abs(s.final_cost - float(g["sol_final_cost"])) <= RTOL_COST * float(g["sol_final_cost"])
    np.testing.assert_allclose([i["cost"] for i in s.iterations], g["sol_iter_cost"], rtol=1e-6)
    assert [i["step_is_successful"] for i in s.iterations] == list(g["sol_iter_success"])
    assert np.abs(prob.poses - g["sol_poses"]).max() < TOL_STATE
    assert np.abs(prob.inv_depth - g["sol_inv_depth"]).max() < TOL_STATE
    if "entry_poses" in g:  # output of the unmodified visnav::bundle_adjustment()
        assert np.abs(prob.poses - g["entry_poses"]).max() < TOL_STATE
        assert np.abs(prob.inv_depth - g["entry_inv_depth"]).max() < TOL_STATE


def test_lm_iterate_is_repeatable_and_matches_minimize_first_step():
    prob, _ = scene(pb.MODE_PHOTOMETRIC, "pinhole", n_kf=10, n_pts=400)
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=9.0))
    a = eng.lm_iterate(1e4)
    b = eng.lm_iterate(1e4)
    assert a == b                                   # state untouched, bit-identical work
    s = eng.minimize()
    it1 = s.iterations[1]
    assert abs(a["cost"] - s.initial_cost) <= 1e-14 * s.initial_cost
    assert abs(a["cost_change"] - it1["cost_change"]) <= 1e-9 * abs(it1["cost_change"])
    assert abs(a["relative_decrease"] - it1["relative_decrease"]) <= 1e-9
    eng.close()


# ------------------------------------------------------------- drop-in proof --
def _dropin_lib():
    import ctypes as C, os
    base = os.path.join(os.path.dirname(of.REF_SO))
    path = os.path.join(base, "libpba_dropin_v4.so" if of.REF_SO.endswith("_v4.so") else "libpba_dropin.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.pba_dropin_solve.argtypes = [C.POINTER(_ffi.pba_problem), C.POINTER(_ffi.pba_options), C.c_int,
                                     C.POINTER(_ffi.pba_summary)]
    return lib


@pytest.mark.parametrize("model", ["pinhole", "ds", "kb4"])
def test_reference_containers_drive_both_solvers(model):
    """The reference's own Corners/Cameras/Landmarks/Calibration containers, built once, are
    optimised (a) by the unmodified visnav::bundle_adjustment() and (b) through
    include/visnav_b200/bundle_adjustment.h -> pba_solve -> CUDA: same argument list, same result."""
    lib = _dropin_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libpba_dropin.so not built on this box")
    prob, _ = scene(pb.MODE_GEOMETRIC, model, n_kf=12, n_pts=600)
    o = of.default_options()
    a, b = prob.copy(), prob.copy()
    s = pb.Summary()
    pa, pbb = a.c, b.c
    assert lib.pba_dropin_solve(C.byref(pa), C.byref(o), 0, None) == 0       # reference, Ceres on the CPU
    assert lib.pba_dropin_solve(C.byref(pbb), C.byref(o), 1, C.byref(s.c)) == 0  # drop-in, CUDA engine
    assert s.gpu_kernel_launches > 0
    assert np.abs(a.poses - prob.poses).max() > 1e-4          # both actually moved
    assert np.abs(b.poses - a.poses).max() < TOL_STATE
    assert np.abs(b.inv_depth - a.inv_depth).max() < TOL_STATE
    fixed = prob.pose_fixed.astype(bool)
    assert np.array_equal(b.poses[fixed], prob.poses[fixed])


@pytest.mark.parametrize("model", ["ds", "kb4"])
def test_config3_shape_lm_matches_oracle(model):
    """BASELINE config-3 shape at a size the CPU oracle finishes in seconds: fisheye model,
    per-frame affine brightness, per-point inverse depth; several BCR levels on the GPU."""
    prob, _ = pb.make_scene(pb.MODE_PHOTOMETRIC, 64, 12000, model)
    po = prob.copy()
    so = of.solve("oracle", po, of.default_options(huber_parameter=9.0, max_num_iterations=8))
    pg = prob.copy()
    sg = pb.bundle_adjustment(pg, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=9.0,
                                                             max_num_iterations=8, solver=pb.SOLVER_BCR))
    assert sg.linear_solver == pb.SOLVER_BCR
    assert sg.num_iterations == so.num_iterations
    assert abs(sg.final_cost - so.final_cost) <= RTOL_COST * so.final_cost
    np.testing.assert_allclose([i["cost"] for i in sg.iterations], [i["cost"] for i in so.iterations], rtol=1e-6)
    assert np.abs(pg.poses - po.poses).max() < TOL_STATE
    assert np.abs(pg.inv_depth - po.inv_depth).max() < TOL_STATE
    assert np.abs(pg.affine - po.affine).max() < 1e-4


def _two_camera_problem(mode, models):
    """Stereo-like rig: keyframes alternate between two calibrations (FrameCamId::cam_id 0 / 1)
    with different camera models, like calib_cam.intrinsics[cam_id] in the reference."""
    prob, _ = scene(mode, models[0], n_kf=10, n_pts=400)
    other, _ = scene(mode, models[1], n_kf=10, n_pts=400)
    intr = np.vstack([prob.intrinsics[0], other.intrinsics[0]])
    cm = np.array([pb._ffi.CAM_NAMES[models[0]], pb._ffi.CAM_NAMES[models[1]]], np.int32)
    pc = (np.arange(prob.n_poses) % 2).astype(np.int32)
    return pb.Problem(mode, prob.poses, prob.pose_fixed, pc, cm, intr, prob.inv_depth, prob.lm_host, prob.lm_host_uv,
                      prob.lm_obs_ptr, prob.obs_target, prob.obs_uv, prob.images, prob.affine)


@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
@pytest.mark.parametrize("models", [("pinhole", "ds"), ("kb4", "eucm")])
def test_mixed_camera_models(mode, models):
    """Two calibrations with different models in one problem (runtime model dispatch).  Geometric
    blocks evaluate the target with the HOST's model name and the target's intrinsic values
    (reprojection.h:97-100, map_utils.h:363-364); photometric blocks use each camera's own model."""
    prob = _two_camera_problem(mode, models)
    hub = huber_for(mode)
    cost_o, r_o, J_o = of.evaluate("oracle", prob, True, hub)
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub))
    cost = eng.evaluate(True)
    assert abs(cost - cost_o) <= 1e-12 * abs(cost_o)
    assert rel(eng.residuals(), r_o) < RTOL_RJ
    assert rel(eng.jacobians(), J_o) < RTOL_RJ
    eng.close()
    if of.have_ref():  # and the oracle agrees with the real reference on this configuration
        cost_r, r_r, J_r = of.evaluate("ref", prob, True, hub)
        assert rel(r_o, r_r) < 1e-11 and rel(J_o, J_r) < 1e-11
