"""Host-side camera layout (pba_analyze_structure: no GPU needed): slots, reverse Cuthill-McKee reordering of
maps whose keyframe order is not banded, RCS block pattern."""
import numpy as np

import pba_b200 as pb


def relabel(prob, perm):
    """The same problem with keyframe i renamed perm[i] (poses / images / affine rows move accordingly)."""
    inv = np.argsort(perm)
    q = prob.copy()
    q.poses = np.ascontiguousarray(prob.poses[inv])
    q.pose_fixed = np.ascontiguousarray(prob.pose_fixed[inv])
    q.pose_calib = np.ascontiguousarray(prob.pose_calib[inv])
    q.lm_host = np.ascontiguousarray(perm[prob.lm_host].astype(np.int32))
    q.obs_target = np.ascontiguousarray(perm[prob.obs_target].astype(np.int32))
    if prob.images is not None:
        q.images = np.ascontiguousarray(prob.images[inv])
    if prob.affine is not None:
        q.affine = np.ascontiguousarray(prob.affine[inv])
    return q


def test_windowed_scene_keeps_its_natural_order():
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 60, 3000, "pinhole")
    slot, ns, bw0, bw1, nb = pb.analyze_structure(prob)
    assert ns == 58 and (slot[:2] == -1).all()              # the first two keyframes are fixed
    assert np.array_equal(slot[2:], np.arange(58))           # natural order kept
    assert bw0 == bw1 and bw1 <= 11
    assert nb >= ns


def test_shuffled_keyframes_are_reordered_to_a_band():
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 120, 6000, "pinhole")
    _, ns, bw_ordered, _, nb_ordered = pb.analyze_structure(prob)
    rng = np.random.default_rng(3)
    perm = rng.permutation(prob.n_poses)
    shuf = relabel(prob, perm)
    slot, ns2, bw0, bw1, nb = pb.analyze_structure(shuf)
    assert ns2 == ns and nb == nb_ordered                    # same graph, same number of blocks
    assert bw0 > 60                                          # the shuffled order is hopeless ...
    assert bw1 <= 2 * bw_ordered                             # ... reverse Cuthill-McKee recovers a narrow band
    free = slot[slot >= 0]
    assert sorted(free.tolist()) == list(range(ns))          # a permutation of the slots


def test_loop_closure_stays_banded():
    """A chain whose last keyframes re-observe landmarks of the first ones: a ring.  Natural half-bandwidth ~ n,
    reverse Cuthill-McKee interleaves the two arms of the ring: ~ twice the window."""
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 200, 8000, "pinhole")
    _, ns, bw_chain, _, _ = pb.analyze_structure(prob)
    q = prob.copy()
    tgt = q.obs_target.copy()
    # landmarks hosted by keyframes 2..5: their last observation moves to keyframes 196..199
    for l in np.nonzero((q.lm_host >= 2) & (q.lm_host <= 5))[0]:
        k = int(q.lm_obs_ptr[l + 1]) - 1
        if k >= int(q.lm_obs_ptr[l]):
            tgt[k] = 196 + (l % 4)
    q.obs_target = np.ascontiguousarray(tgt)
    slot, ns2, bw0, bw1, _ = pb.analyze_structure(q)
    assert ns2 == ns
    assert bw0 >= 190
    assert bw1 <= 2 * bw_chain + 4


def test_grid_flight_is_not_banded_under_any_order():
    """Lawn-mower flight (make_grid_scene): covisibility is a 2-D mesh, so reverse Cuthill-McKee cannot do better
    than ~ one flight line of keyframes; the layout keeps the narrower of the two orders and reports it."""
    prob, _ = pb.make_grid_scene(12, 12, 3000)
    slot, ns, bw0, bw1, nb = pb.analyze_structure(prob)
    assert ns == 142 and bw1 <= bw0
    assert bw1 >= 12                                          # wider than the banded solvers take
    assert nb < ns * (ns + 1) // 2                            # but far from dense


def test_grid_flight_oracle_matches_the_reference_golden():
    """The CPU restatement on the non-banded 27 x 27 flight (4,362 unknowns) against the real reference's run
    (tests/golden/scale_grid.npz, made by make_golden_scale.py grid)."""
    import os
    import oracle_ffi as of
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_grid.npz"))
    prob, _ = pb.make_grid_scene(27, 27, int(g["n_pts"]))
    assert prob.n_obs == int(g["n_obs"])
    s = of.solve("oracle", prob, of.default_options(huber_parameter=1.0, max_num_iterations=20))
    cost = np.array([i["cost"] for i in s.iterations])
    assert len(cost) == len(g["iter_cost"])
    assert np.all(np.abs(cost - g["iter_cost"]) <= 1e-9 * np.abs(g["iter_cost"]))
    assert np.abs(prob.poses - g["sol_poses"]).max() < 1e-9


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
def test_shuffled_keyframes_solve_with_an_exact_banded_solver(mode):
    """An 'unordered map': the keyframes of a windowed scene in random order.  The engine renumbers the cameras
    (reverse Cuthill-McKee), so AUTO still picks an exact banded solver, and the LM run matches the oracle."""
    import oracle_ffi as of
    hub = 9.0 if mode == pb.MODE_PHOTOMETRIC else 1.0
    prob, _ = pb.make_scene(mode, 90, 2500, "pinhole")
    rng = np.random.default_rng(11)
    perm = rng.permutation(prob.n_poses)
    shuf = relabel(prob, perm)
    po, pg = shuf.copy(), shuf.copy()
    so = of.solve("oracle", po, of.default_options(huber_parameter=hub, max_num_iterations=8))
    sg = pb.bundle_adjustment(pg, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, max_num_iterations=8))
    assert sg.linear_solver in (pb.SOLVER_BAND, pb.SOLVER_BCR)
    assert sg.num_inexact_linear_solves == 0
    assert sg.num_iterations == so.num_iterations
    assert abs(sg.final_cost - so.final_cost) <= 1e-6 * so.final_cost
    assert np.abs(pg.poses - po.poses).max() < 1e-5
    assert np.abs(pg.inv_depth - po.inv_depth).max() < 1e-5


def test_bad_observation_indices_are_rejected():
    """The per-observation part of the input validation rides on the layout's observation scan (host.cu
    analyze_cameras): an index out of range or a landmark observed by its own host is PBA_ERR_INVALID_ARGUMENT,
    as the reference's containers make them impossible (map_utils.h:340-376 iterates obs of existing cameras)."""
    import pytest
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 12, 400, "pinhole")
    pb.analyze_structure(prob)
    k = int(prob.lm_obs_ptr[7])
    for field, idx, val in (("obs_target", k, prob.n_poses), ("obs_target", k, -1),
                            ("obs_target", k, int(prob.lm_host[7])), ("lm_host", 7, prob.n_poses),
                            ("lm_host", 399, -3)):
        bad = prob.copy()
        arr = getattr(bad, field).copy()
        arr[idx] = val
        setattr(bad, field, arr)
        with pytest.raises(RuntimeError, match="INVALID_ARGUMENT"):
            pb.analyze_structure(bad)
    bad = prob.copy()
    ptr = bad.lm_obs_ptr.copy()
    ptr[5] = ptr[6] + 1  # not monotone
    bad.lm_obs_ptr = ptr
    with pytest.raises(RuntimeError, match="INVALID_ARGUMENT"):
        pb.analyze_structure(bad)
