"""Multi-rank path: landmark sharding (host logic, CPU/gloo, world_size 2) and the
sharded LM solve with the NCCL all-reduce of the partial RCS (needs >= 2 GPUs)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import oracle_ffi as of
import pba_b200 as pb

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "multi_rank_worker.py")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_landmarks_properties():
    rng = np.random.default_rng(0)
    for n_lm, world in ((1, 1), (7, 2), (1000, 4), (1000, 8), (5, 8)):
        cnt = rng.integers(0, 12, size=n_lm)
        ptr = np.r_[0, np.cumsum(cnt)].astype(np.int64)
        b = pb.partition_landmarks(ptr, world)
        assert b[0] == 0 and b[-1] == n_lm and len(b) == world + 1
        assert all(b[i] <= b[i + 1] for i in range(world))
        if n_lm >= 100:  # balanced by observation count
            loads = [ptr[b[i + 1]] - ptr[b[i]] for i in range(world)]
            assert max(loads) - min(loads) <= 24


def test_two_rank_sharding_gloo(tmp_path):
    """world_size 2 over gloo: shard costs / pose gradients all-reduce to the whole problem's."""
    port = free_port()
    procs = [subprocess.Popen([sys.executable, WORKER, "gloo", str(r), "2", str(tmp_path), str(pb.MODE_GEOMETRIC),
                               str(port)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    out = json.load(open(tmp_path / "gloo.json"))
    prob, _ = pb.make_scene(pb.MODE_GEOMETRIC, 14, 900, "pinhole")
    cost, r, J = of.evaluate("oracle", prob, True, 1.0, threads=2)
    assert out["n_obs"] == prob.n_obs and out["n_lm"] == prob.n_landmarks
    assert abs(out["cost"] - cost) <= 1e-12 * cost
    g = np.zeros((prob.n_poses, 6))
    for l in range(prob.n_landmarks):
        for k in range(int(prob.lm_obs_ptr[l]), int(prob.lm_obs_ptr[l + 1])):
            g[prob.lm_host[l]] += J[k][:, 0:6].T @ r[k]
            g[prob.obs_target[k]] += J[k][:, 6:12].T @ r[k]
    np.testing.assert_allclose(np.load(tmp_path / "gloo_grad.npy").reshape(-1, 6), g, rtol=1e-11, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("collective", ["peer", "nccl"])
@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
def test_two_gpu_sharded_lm_matches_single_gpu(tmp_path, mode, collective):
    """One process per GPU.  `peer`: the library's own NVLink collectives over cudaIpc-mapped buffers (csrc/peer.cu,
    the default); `nccl`: PBA_NO_PEER=1 keeps ncclAllReduce.  Both must reproduce the one-GPU solve."""
    if pb.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ)
    env.pop("PBA_NO_PEER", None)
    if collective == "nccl":
        env["PBA_NO_PEER"] = "1"
    procs = [subprocess.Popen([sys.executable, WORKER, "gpu", str(r), "2", str(tmp_path), str(mode)], env=env)
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    prob, _ = pb.make_scene(mode, 14, 900, "pinhole")
    hub = 9.0 if mode == pb.MODE_PHOTOMETRIC else 1.0
    ref = prob.copy()
    s = pb.bundle_adjustment(ref, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub))
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert str(r0["collective"]) == str(r1["collective"])  # the choice is collective
    if collective == "nccl":
        assert str(r0["collective"]) == "nccl"
    for r in (r0, r1):  # every rank reports the global cost trace
        assert int(r["iterations"]) == s.num_iterations
        assert abs(float(r["cost0"]) - s.initial_cost) <= 1e-12 * s.initial_cost
        assert abs(float(r["final_cost"]) - s.final_cost) <= 1e-9 * s.final_cost
        np.testing.assert_allclose(r["costs"], [i["cost"] for i in s.iterations], rtol=1e-9)
        assert np.abs(r["poses"] - ref.poses).max() < 1e-8
    rho = np.r_[r0["rho"], r1["rho"]]
    assert int(r0["first"]) == 0 and int(r1["first"]) == len(r0["rho"])
    assert np.abs(rho - ref.inv_depth).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
def test_single_process_two_gpus_matches_single_gpu(mode):
    """The drop-in call from ONE process over two GPUs (pba_options.num_gpus; the reference's caller is a single
    thread, src/sfm.cpp:1903-1913): same iteration trace and state as the one-GPU solve."""
    if pb.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    hub = 9.0 if mode == pb.MODE_PHOTOMETRIC else 1.0
    prob, _ = pb.make_scene(mode, 14, 900, "pinhole")
    one, two = prob.copy(), prob.copy()
    s1 = pb.bundle_adjustment(one, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub))
    pb.multi_gpu_init(0, 2)
    s2 = pb.bundle_adjustment(two, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, num_gpus=2))
    assert s2.num_iterations == s1.num_iterations and s2.termination_type == s1.termination_type
    np.testing.assert_allclose([i["cost"] for i in s2.iterations], [i["cost"] for i in s1.iterations], rtol=1e-9)
    assert abs(s2.final_cost - s1.final_cost) <= 1e-9 * s1.final_cost
    assert np.abs(two.poses - one.poses).max() < 1e-8
    assert np.abs(two.inv_depth - one.inv_depth).max() < 1e-8
    if mode == pb.MODE_PHOTOMETRIC:
        assert np.abs(two.affine - one.affine).max() < 1e-8
    # a second call reuses the cached communicators and the device arena
    again = prob.copy()
    s3 = pb.bundle_adjustment(again, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, num_gpus=2))
    assert s3.final_cost == s2.final_cost and np.array_equal(again.poses, two.poses)


@pytest.mark.gpu
def test_second_device_after_first():
    """Handles on two devices in one process (the shared-memory opt-in is per device)."""
    if pb.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    prob, _ = pb.make_scene(pb.MODE_PHOTOMETRIC, 14, 900, "pinhole")
    out = []
    for dev in (0, 1, 0):
        p = prob.copy()
        out.append(pb.bundle_adjustment(p, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=9.0, device=dev)))
    assert out[0].final_cost == out[1].final_cost == out[2].final_cost


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [pb.MODE_GEOMETRIC, pb.MODE_PHOTOMETRIC])
def test_shards_evaluated_on_one_gpu_tile_the_whole_problem(mode):
    """Runs on a ONE-GPU box (the 2-GPU tests above are skipped there): the three landmark shards of a problem,
    created with (rank, 3) on the same device and evaluated without any collective, hold exactly the residual and
    Jacobian blocks of the whole problem's engine — same partition rule as partition_landmarks, same block order,
    bit-identical values (a block's arithmetic does not depend on which shard it lives in)."""
    hub = 9.0 if mode == pb.MODE_PHOTOMETRIC else 1.0
    prob, _ = pb.make_scene(mode, 14, 900, "pinhole")
    opts = pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub)
    whole = pb.Engine(prob, opts)
    whole.evaluate(True)
    r, J = whole.residuals(), whole.jacobians()
    whole.close()
    world = 3
    b = pb.partition_landmarks(prob.lm_obs_ptr, world)
    covered = 0
    for rank in range(world):
        eng = pb.Engine(prob, opts, rank=rank, world_size=world)
        o0, o1 = int(prob.lm_obs_ptr[b[rank]]), int(prob.lm_obs_ptr[b[rank + 1]])
        assert eng.first_landmark == b[rank] and eng.n_landmarks_local == b[rank + 1] - b[rank]
        assert eng.n_obs_local == o1 - o0
        eng.evaluate(True, want_cost=False)          # a cost would need the all-reduce; the blocks do not
        rs, Js = eng.residuals(), eng.jacobians()
        eng.close()
        assert np.array_equal(rs, r[o0:o1])
        assert np.array_equal(Js, J[o0:o1])
        covered += o1 - o0
    assert covered == prob.n_obs


@pytest.mark.gpu
@pytest.mark.parametrize("mode,ranks", [(pb.MODE_GEOMETRIC, 2), (pb.MODE_PHOTOMETRIC, 2), (pb.MODE_PHOTOMETRIC, 3),
                                        (pb.MODE_GEOMETRIC, 8)])
def test_sharded_solve_with_ranks_emulated_on_one_gpu(mode, ranks, monkeypatch):
    """The whole sharded LM solve on a ONE-GPU box: PBA_EMULATE_RANKS=1 puts every rank's handle on the same device
    (one host thread per rank, as the single-process multi-GPU solve does) and replaces only the transport — a host
    barrier + one sum kernel instead of the NVLink / NCCL exchange.  Partition, per-rank layouts, the packed payload,
    scaling after the reduction, the redundant RCS solve, per-shard back-substitution and the scalar exchange are the
    production code; the result must equal the one-rank solve like the real 2-GPU run does."""
    hub = 9.0 if mode == pb.MODE_PHOTOMETRIC else 1.0
    prob, _ = pb.make_scene(mode, 14, 900, "pinhole")
    one, many = prob.copy(), prob.copy()
    s1 = pb.bundle_adjustment(one, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub))
    monkeypatch.setenv("PBA_EMULATE_RANKS", "1")
    s2 = pb.bundle_adjustment(many, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, num_gpus=ranks))
    assert s2.num_iterations == s1.num_iterations and s2.termination_type == s1.termination_type
    assert [i["step_is_successful"] for i in s2.iterations] == [i["step_is_successful"] for i in s1.iterations]
    np.testing.assert_allclose([i["cost"] for i in s2.iterations], [i["cost"] for i in s1.iterations], rtol=1e-9)
    np.testing.assert_allclose([i["gradient_max_norm"] for i in s2.iterations],
                               [i["gradient_max_norm"] for i in s1.iterations], rtol=1e-6)
    assert abs(s2.final_cost - s1.final_cost) <= 1e-9 * s1.final_cost
    assert np.abs(many.poses - one.poses).max() < 1e-8
    assert np.abs(many.inv_depth - one.inv_depth).max() < 1e-8
    if mode == pb.MODE_PHOTOMETRIC:
        assert np.abs(many.affine - one.affine).max() < 1e-8
