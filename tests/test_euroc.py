"""BASELINE config 1: bundle adjustment on the bundled data/euroc_V1 keyframes.

tools/euroc/calibrate.py + tools/euroc/build_map.py (offline, in the build container) turn the reference's EuRoC
sample — 82 stereo keyframes seconds apart, a loop-y room trajectory — into a real map: stereo calibration through
the reference's own calibration functor, cv2 keypoints / matches, feature tracks, PnP localisation, stereo
triangulation, the reference's bundle_adjustment() every few cameras.  The fixtures hold
  euroc_v1_map.npz    152 cameras, 3,930 landmarks, 15,868 residual blocks, double-sphere model; the reduced camera
                      system is NOT banded (half-bandwidth 149 of 150 slots, 113 after reverse Cuthill-McKee); the
                      state before the last optimisation + the result of the reference's own bundle_adjustment();
  euroc_v1_photo.npz  the photometric problem on the first 12 keyframes with their real images + the reference
                      functor's residuals / Jacobians and LM run (vendored Ceres AutoDiff).
CPU: the oracle restatement against both.  GPU: the CUDA engine through the C ABI against both.
"""
import os

import numpy as np
import pytest

import oracle_ffi as of
import pba_b200 as pb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_geom():
    g = np.load(os.path.join(GOLDEN, "euroc_v1_map.npz"))
    prob = pb.Problem(int(g["mode"]), g["poses"], g["pose_fixed"], g["pose_calib"], g["calib_model"], g["intrinsics"],
                      g["inv_depth"], g["lm_host"], g["lm_host_uv"], g["lm_obs_ptr"], g["obs_target"], g["obs_uv"])
    return g, prob


def load_photo():
    g = np.load(os.path.join(GOLDEN, "euroc_v1_photo.npz"))
    prob = pb.Problem(int(g["mode"]), g["poses"], g["pose_fixed"], g["pose_calib"], g["calib_model"], g["intrinsics"],
                      g["inv_depth"], g["lm_host"], g["lm_host_uv"], g["lm_obs_ptr"], g["obs_target"], None, g["images"],
                      g["affine"])
    return g, prob


def check_lm(g, prob, s, with_affine=False):
    assert s.num_iterations == len(g["sol_iter_cost"])
    assert [int(i["step_is_successful"]) for i in s.iterations] == [int(x) for x in g["sol_iter_success"]]
    cost = np.array([i["cost"] for i in s.iterations])
    assert np.all(np.abs(cost - g["sol_iter_cost"]) <= 1e-6 * g["sol_iter_cost"])
    assert abs(s.final_cost - float(g["sol_final_cost"])) <= 1e-6 * float(g["sol_final_cost"])
    assert s.termination_type == int(g["sol_termination"])
    assert np.abs(prob.poses - g["sol_poses"]).max() < 1e-5
    assert np.abs(prob.inv_depth - g["sol_inv_depth"]).max() < 1e-5
    if with_affine:
        assert np.abs(prob.affine - g["sol_affine"]).max() < 1e-5 * max(1.0, np.abs(g["sol_affine"]).max())


def test_calibration_of_the_euroc_rig():
    """tools/euroc/opt_calib.json: what the reference's calibration application would write (same functor, same data)."""
    cal = pb.load_calibration(os.path.join(ROOT, "tools", "euroc", "opt_calib.json"))
    assert cal.models == ["ds", "ds"]
    assert abs(np.linalg.norm(cal.T_i_c[1][4:]) - 0.110) < 0.002      # the EuRoC stereo baseline
    assert np.all(np.abs(cal.intrinsics[:, 0] - 356) < 12) and np.all(np.abs(cal.intrinsics[:, 5] - 0.57) < 0.05)


def test_euroc_map_is_a_real_non_banded_problem():
    g, prob = load_geom()
    assert prob.n_poses == 152 and prob.n_obs > 15000 and prob.n_landmarks > 3500
    slot, ns, bw0, bw1, nb = pb.analyze_structure(prob)
    assert ns == 150 and (slot[prob.pose_fixed == 1] == -1).all()
    assert bw0 > 140 and bw1 < bw0 and bw1 > 60          # loop-y room trajectory: no ordering makes it narrow
    # the reference's unmodified entry point and the harness solve agree (same Ceres, same functor)
    assert np.abs(g["entry_poses"] - g["sol_poses"]).max() < 1e-9


def test_oracle_matches_reference_on_the_euroc_map():
    g, prob = load_geom()
    s = of.solve("oracle", prob, of.default_options(huber_parameter=float(g["huber"])))
    check_lm(g, prob, s)


def test_oracle_matches_reference_on_the_euroc_images():
    g, prob = load_photo()
    cost, r, J = of.evaluate("oracle", prob, True, float(g["huber"]))
    sel = g["ref_sel"]
    assert abs(cost - float(g["ref_cost"])) <= 1e-12 * float(g["ref_cost"])
    assert np.abs(r[sel] - g["ref_residuals"]).max() <= 1e-11 * np.abs(g["ref_residuals"]).max()
    assert np.abs(J[sel] - g["ref_jacobians"]).max() <= 1e-11 * np.abs(g["ref_jacobians"]).max()
    s = of.solve("oracle", prob, of.default_options(huber_parameter=float(g["huber"]),
                                                    max_num_iterations=int(g["max_num_iterations"])))
    check_lm(g, prob, s, with_affine=True)


@pytest.mark.gpu
def test_cuda_matches_reference_on_the_euroc_map():
    g, prob = load_geom()
    s = pb.bundle_adjustment(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=float(g["huber"])))
    assert s.gpu_kernel_launches > 0
    assert s.linear_solver == pb.SOLVER_CHOLESKY and s.num_inexact_linear_solves == 0   # non-banded: exact dense solve
    check_lm(g, prob, s)
    assert np.abs(prob.poses - g["entry_poses"]).max() < 1e-5           # = the unmodified visnav::bundle_adjustment()
    assert np.abs(prob.inv_depth - g["entry_inv_depth"]).max() < 1e-5


@pytest.mark.gpu
def test_cuda_matches_reference_on_the_euroc_images():
    g, prob = load_photo()
    eng = pb.Engine(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=float(g["huber"])))
    cost = eng.evaluate(True)
    r, J = eng.blocks(g["ref_sel"])
    eng.close()
    assert abs(cost - float(g["ref_cost"])) <= 1e-10 * float(g["ref_cost"])
    assert np.abs(r - g["ref_residuals"]).max() <= 1e-9 * np.abs(g["ref_residuals"]).max()
    assert np.abs(J - g["ref_jacobians"]).max() <= 1e-9 * np.abs(g["ref_jacobians"]).max()
    s = pb.bundle_adjustment(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=float(g["huber"]),
                                                              max_num_iterations=int(g["max_num_iterations"])))
    check_lm(g, prob, s, with_affine=True)
