"""Generates tests/golden/scale_*.npz: LM runs of the REAL reference (oracle/_ref:
unmodified visnav functor + vendored Ceres 2.0.0, SPARSE_SCHUR, options of
include/visnav/map_utils.h:378-383) on BASELINE.json's configurations AT THEIR
STATED SIZES.  The scenes are the deterministic synthetic ones of SURVEY.md §8(d)
(pba_b200.make_scene, CPU renderer), so a fixture only has to hold the outputs:
per-iteration cost / radius / accept flags, the final poses and affine terms and a
strided sample of the final inverse distances.

    python tests/golden/make_golden_scale.py cfg2 cfg3_ds cfg3_kb4 cfg5 cfg4

cfg4 (2,000 KF x 2M points, 18M residual blocks; run on the 99.6 % keyframe prefix the
vendored Ceres can represent) needs ~35 GB of RAM and ~10 min on 8 cores for its 3 LM iterations; cfg5 (1,000 cameras x 1M landmarks) ~8 min for 20.
tests/test_gpu_scale.py compares the CUDA engine against these on the GPU box, where
/root/reference does not exist.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_ffi as of  # noqa: E402
import pba_b200 as pb  # noqa: E402

# name -> (mode, keyframes, points, camera model, max LM iterations)
CASES = {
    "cfg2": (pb.MODE_PHOTOMETRIC, 50, 20000, "pinhole", 20),
    "cfg3_ds": (pb.MODE_PHOTOMETRIC, 200, 100000, "ds", 20),
    "cfg3_kb4": (pb.MODE_PHOTOMETRIC, 200, 100000, "kb4", 20),
    "cfg4": (pb.MODE_PHOTOMETRIC, 2000, 2000000, "pinhole", 3),
    "cfg5": (pb.MODE_GEOMETRIC, 1000, 1000000, "pinhole", 20),
    # NOT banded under any camera order: 27 x 27 keyframes of a lawn-mower flight, covisible along
    # both grid directions; 727 free cameras = 4,362 unknowns, half-bandwidth 113 cameras after reverse Cuthill-McKee
    "grid": (pb.MODE_GEOMETRIC, 729, 30000, "pinhole", 20),
}
GRID = (27, 27)
RHO_STRIDE = 97  # final inverse distances kept: every 97th landmark


def main(names):
    assert of.have_ref(), "build oracle/_ref first (make ref)"
    for name in names:
        mode, kf, pts, model, iters = CASES[name]
        hub = 9.0 if mode == pb.MODE_PHOTOMETRIC else 1.0
        t0 = time.time()
        prob, _ = pb.make_grid_scene(GRID[0], GRID[1], pts) if name == "grid" else pb.make_scene(mode, kf, pts, model)
        # cfg4: the vendored Ceres aborts above 2^31 - 1 Jacobian entries (oracle_ffi.CERES_MAX_NONZEROS); the
        # fixture is made on the longest keyframe prefix it can hold (99.6 % of the observations) and the GPU
        # test solves that same prefix
        n_pref, prob = of.largest_reference_prefix(prob)
        t1 = time.time()
        s = of.solve("ref", prob, of.default_options(huber_parameter=hub, max_num_iterations=iters))
        t2 = time.time()
        its = s.iterations
        out = dict(
            mode=mode, n_kf=kf, n_pts=pts, n_kf_prefix=n_pref, model=pb._ffi.CAM_NAMES[model], huber=hub, max_num_iterations=iters,
            n_obs=prob.n_obs, rho_stride=RHO_STRIDE,
            initial_cost=s.initial_cost, final_cost=s.final_cost, termination=s.termination_type,
            iter_cost=np.array([i["cost"] for i in its]),
            iter_radius=np.array([i["trust_region_radius"] for i in its]),
            iter_success=np.array([i["step_is_successful"] for i in its]),
            iter_gradient_max_norm=np.array([i["gradient_max_norm"] for i in its]),
            sol_poses=prob.poses, sol_inv_depth_sample=prob.inv_depth[::RHO_STRIDE].copy(),
            sol_inv_depth_sum=float(prob.inv_depth.sum()),
            ref_threads=os.cpu_count(), ref_minimizer_s=s.minimizer_time_in_seconds,
            ref_jacobian_s=s.jacobian_evaluation_time_in_seconds, ref_jacobian_evals=s.num_jacobian_evaluations,
            ref_linear_solver_s=s.linear_solver_time_in_seconds, ref_total_s=t2 - t1,
        )
        if mode == pb.MODE_PHOTOMETRIC:
            out["sol_affine"] = prob.affine
        np.savez_compressed(os.path.join(HERE, "scale_%s.npz" % name), **out)
        print("%s: %d obs, scene %.1f s, reference %.1f s (minimizer %.1f), %d iterations, cost %.9e -> %.9e, %s"
              % (name, prob.n_obs, t1 - t0, t2 - t1, s.minimizer_time_in_seconds, len(its), s.initial_cost,
                 s.final_cost, s.message), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["cfg2", "cfg3_ds", "cfg3_kb4"])
