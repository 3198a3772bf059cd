"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref/libpba_ref.so:
the unmodified visnav headers + vendored Ceres 2.0.0, built by oracle/ref/Makefile).

Run in the build container (needs /root/reference to have been compiled):
    python tests/golden/make_golden.py
Each fixture holds the flat problem (inputs), the reference's robustified
per-block residuals and local Jacobians (ceres::Problem::Evaluate), and the
reference's full LM trace + final state (ceres::Solve with the options of
include/visnav/map_utils.h:378-383).  The geometric fixtures' final state comes
from the unmodified visnav::bundle_adjustment() entry point.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_ffi as of  # noqa: E402
import pba_b200 as pb  # noqa: E402

W, H = 192, 144
SMALL_INTR = {
    "pinhole": [95.0, 95.0, 95.5, 71.5, 0, 0, 0, 0],
    "ds": [95.0, 95.0, 95.5, 71.5, -0.2, 0.55, 0, 0],
    "kb4": [97.0, 97.0, 95.5, 71.5, 0.00693023, -0.0013828, -0.000272596, -0.000452646],
    "eucm": [95.0, 95.0, 95.5, 71.5, 0.55, 1.05, 0, 0],
}
CASES = [(mode, model) for mode in (pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC) for model in ("pinhole", "ds", "kb4", "eucm")]


def make_case(mode, model):
    import ctypes as C
    intr = (C.c_double * 8)(*SMALL_INTR[model])
    over = dict(intrinsics=intr, min_len=4, max_len=6)
    if mode == pb.MODE_PHOTOMETRIC:
        over.update(pose_sigma=0.0015)
    prob, gt = pb.make_scene(mode, 7, 70, model, width=W, height=H, **over)
    return prob, gt


def main():
    assert of.have_ref(), "build oracle/_ref first (make ref)"
    for mode, model in CASES:
        prob, gt = make_case(mode, model)
        hub = 9.0 if mode == pb.MODE_PHOTOMETRIC else 1.0
        cost, r, J = of.evaluate("ref", prob, True, hub)
        cost_nh, r_nh, J_nh = of.evaluate("ref", prob, False, hub)
        sol = prob.copy()
        s = of.solve("ref", sol, of.default_options(huber_parameter=hub))
        out = dict(
            mode=mode, model=pb._ffi.CAM_NAMES[model], huber=hub,
            poses=prob.poses, pose_fixed=prob.pose_fixed, pose_calib=prob.pose_calib, calib_model=prob.calib_model,
            intrinsics=prob.intrinsics, inv_depth=prob.inv_depth, lm_host=prob.lm_host, lm_host_uv=prob.lm_host_uv,
            lm_obs_ptr=prob.lm_obs_ptr, obs_target=prob.obs_target,
            ref_cost=cost, ref_residuals=r, ref_jacobians=J, ref_cost_nohuber=cost_nh, ref_residuals_nohuber=r_nh,
            ref_jacobians_nohuber=J_nh,
            sol_poses=sol.poses, sol_inv_depth=sol.inv_depth,
            sol_initial_cost=s.initial_cost, sol_final_cost=s.final_cost, sol_termination=s.termination_type,
            sol_iter_cost=np.array([i["cost"] for i in s.iterations]),
            sol_iter_radius=np.array([i["trust_region_radius"] for i in s.iterations]),
            sol_iter_success=np.array([i["step_is_successful"] for i in s.iterations]),
            sol_iter_step_norm=np.array([i["step_norm"] for i in s.iterations]),
            sol_iter_gradient_max_norm=np.array([i["gradient_max_norm"] for i in s.iterations]),
        )
        if mode == pb.MODE_GEOMETRIC:
            out["obs_uv"] = prob.obs_uv
            entry = prob.copy()
            of.solve("ref", entry, of.default_options(huber_parameter=hub), use_reference_entry=True)
            out["entry_poses"] = entry.poses          # unmodified visnav::bundle_adjustment()
            out["entry_inv_depth"] = entry.inv_depth
        else:
            out["images"] = prob.images
            out["affine"] = prob.affine
            out["sol_affine"] = sol.affine
        name = "%s_%s.npz" % ("photo" if mode else "geom", model)
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, "obs", prob.n_obs, "cost %.6e -> %.6e in %d iterations" % (s.initial_cost, s.final_cost, s.num_iterations))


def perturb_for_outliers(prob, seed):
    """Push a few landmarks into each outlier class of src/sfm.cpp:1928-1952."""
    rng = np.random.default_rng(seed)
    idx = rng.permutation(prob.n_landmarks)
    prob.inv_depth[idx[0:4]] *= 60.0     # a few cm from the host camera -> camera-distance / z flags
    prob.inv_depth[idx[4:8]] *= -1.0     # behind the host -> z flag, huge reprojection error elsewhere
    prob.inv_depth[idx[8:14]] *= 1.6     # moderate depth error -> "normal" reprojection error
    prob._c = None
    return prob


def main_projections():
    """projections_<model>.npz: the reference's Landmark::get_p / SE3::inverse /
    AbstractCamera::project on its own containers (pba_ref_compute_projections),
    at the default thresholds of src/sfm.cpp:254-261."""
    assert of.have_ref(), "build oracle/_ref first (make ref)"
    for k, model in enumerate(("pinhole", "ds", "kb4", "eucm")):
        prob, _ = make_case(pb.MODE_GEOMETRIC, model)
        prob = perturb_for_outliers(prob, 100 + k)
        out = of.compute_projections("ref", prob)
        p_w = of.landmark_positions("ref", prob)
        name = "projections_%s.npz" % model
        np.savez_compressed(
            os.path.join(HERE, name), mode=prob.mode, poses=prob.poses, pose_fixed=prob.pose_fixed,
            pose_calib=prob.pose_calib, calib_model=prob.calib_model, intrinsics=prob.intrinsics,
            inv_depth=prob.inv_depth, lm_host=prob.lm_host, lm_host_uv=prob.lm_host_uv, lm_obs_ptr=prob.lm_obs_ptr,
            obs_target=prob.obs_target, obs_uv=prob.obs_uv, ref_p_w=p_w, **{"ref_" + k2: v for k2, v in out.items()})
        fl = out["outlier_flags"]
        print(name, "slots", fl.size, "flag counts", [int(np.sum((fl >> b) & 1)) for b in range(4)])


def main_map():
    """map_geom_ds.cereal: the reference's own save_map_file (include/visnav/map_utils.h:58-86, cereal
    binary) on containers built from the geom_ds fixture (+ deterministic filler, oracle/ref/ref_harness.cpp)."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as gu
    assert of.have_ref(), "build oracle/_ref first (make ref)"
    prob, _ = gu.load(os.path.join(HERE, "geom_ds.npz"))
    lib = of.ref()
    lib.pba_ref_save_map.argtypes = [C.POINTER(pb._ffi.pba_problem), C.c_char_p]
    pc = prob.c
    out = os.path.join(HERE, "map_geom_ds.cereal")
    assert lib.pba_ref_save_map(C.byref(pc), out.encode()) == 0
    print("map_geom_ds.cereal", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    if "--map" in sys.argv:
        main_map()
    elif "--projections" in sys.argv:
        main_projections()
    else:
        main()
