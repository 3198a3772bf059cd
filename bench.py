#!/usr/bin/env python
"""bench.py — LM iterations/s and residual+Jacobian evaluations/s of the B200
bundle-adjustment engine on BASELINE.json's headline workload (synthetic
photometric BA, 2,000 keyframes x 2M points, ~18M observations), next to the
reference's own CPU path (vendored Ceres + visnav functor) on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU)

A "step" is one complete Levenberg-Marquardt iteration at the perturbed initial
state (pba_lm_iterate): one residual+Jacobian evaluation (K1), normal equations
+ Schur elimination (+ NCCL all-reduce of the partial RCS when N > 1), RCS
solve, back-substitution, model cost, candidate point and its cost (K2) — the
work of one successful Ceres iteration.  The state is not advanced, so every
step does identical work.

Rank 0 prints ONE JSON line (see the task contract): `value` = LM iterations/s
with inputs resident in HBM; `e2e` = the same metric through the drop-in call
with HOST buffers (flatten + H2D + solve + D2H inside the timed region);
`roofline` = K1's algorithmic bytes / CUDA-event time against the measured HBM
peak; `cpu_baseline` = the reference on a bounded sample of the same scene.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# Algorithmic bytes per observation of the residual/Jacobian kernel.  SURVEY.md §8(d) counts 1,088 B for a
# photometric kernel that MATERIALISES the 8 x 15 local Jacobian (inputs 64 + residual 64 + J 960).  K1 does not
# store the six target-pose columns (they are host-pose columns x a per-edge 6x6 adjoint) but it
# does emit the 128 B Schur record (E^T [J r], so that no later kernel re-reads J for the elimination), so per
# §8(d) it is reported against its OWN compulsory bytes: 1,088 - 8 rows x 6 columns x 8 B + 128 B = 832
# (ncu DRAM traffic: 875 B/obs, profiles/k1_traffic.json).  The figure without the Schur record (704) and the
# materialised one (1,088) are kept next to it for comparability.
# (geometric: 256 - 2 rows x 6 columns x 8 B = 160, + 128 B Schur record = 288)
K1_BYTES_PER_OBS = {1: 832, 0: 288}
K1_BYTES_PER_OBS_JR_ONLY = {1: 704, 0: 160}
K1_BYTES_PER_OBS_MATERIALISED = {1: 1088, 0: 256}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kf", type=int, default=2000)
    ap.add_argument("--pts", type=int, default=2000000)
    ap.add_argument("--model", default="pinhole")
    ap.add_argument("--mode", type=int, default=1, help="1 photometric (headline), 0 geometric")
    ap.add_argument("--solver", type=int, default=0, help="0 auto, 1 cholesky, 2 pcg")
    ap.add_argument("--sample-kf", type=int, default=48, help="keyframes in the CPU-baseline sample")
    ap.add_argument("--cpu-iters", type=int, default=3, help="LM iterations of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a, n_obs):
    kind = "photometric (8-px pattern, affine brightness, Huber 9)" if a.mode == 1 else "geometric reprojection (Huber 1)"
    return "%s BA, %d KF x %d pts (%d obs), %s 752x480, synthetic textured wall" % (kind, a.kf, a.pts, n_obs, a.model)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = "/tmp/pba_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    out["sm_max_mhz"] = float(p[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def sample_problem(prob, n_kf_s):
    """Bounded sample of the same scene: the first n_kf_s keyframes and every
    landmark whose whole track lies inside them."""
    import pba_b200 as pb
    last_target = np.maximum.reduceat(prob.obs_target, prob.lm_obs_ptr[:-1].clip(max=prob.n_obs - 1))
    inside = last_target < n_kf_s
    L = int(np.argmin(inside)) if not inside.all() else prob.n_landmarks
    o1 = int(prob.lm_obs_ptr[L])
    return pb.Problem(prob.mode, prob.poses[:n_kf_s].copy(), prob.pose_fixed[:n_kf_s], prob.pose_calib[:n_kf_s],
                      prob.calib_model, prob.intrinsics, prob.inv_depth[:L].copy(), prob.lm_host[:L],
                      prob.lm_host_uv[:L], prob.lm_obs_ptr[:L + 1], prob.obs_target[:o1],
                      None if prob.obs_uv is None else prob.obs_uv[:o1],
                      None if prob.images is None else prob.images[:n_kf_s],
                      None if prob.affine is None else prob.affine[:n_kf_s].copy())


def run_cpu_reference(prob, a, iters, n_obs_full):
    """The reference's CPU implementation (oracle/_ref when built, else the
    oracle port) on a bounded sample, all host threads.  Returns the
    cpu_baseline dict; LM it/s is also given scaled to the full workload by the
    observation ratio (linear-cost assumption, optimistic for the CPU)."""
    import oracle_ffi as of
    kind = "reference" if of.have_ref() else "port"
    lib = "ref" if kind == "reference" else "oracle"
    cores = os.cpu_count() or 1
    hub = 9.0 if a.mode == 1 else 1.0
    smp = sample_problem(prob, min(a.sample_kf, prob.n_poses))
    opts = of.default_options(huber_parameter=hub, max_num_iterations=iters)
    t0 = time.time()
    s = of.solve(lib, smp, opts, threads=cores)
    wall = time.time() - t0
    lm_its = max(s.num_iterations - 1, 1)
    it_per_s = lm_its / max(s.minimizer_time_in_seconds, 1e-9)
    resjac = smp.n_obs * s.num_jacobian_evaluations / max(s.jacobian_evaluation_time_in_seconds, 1e-9)
    scale = smp.n_obs / float(n_obs_full)
    return {
        "value": it_per_s * scale, "unit": "LM it/s", "cores": cores, "kind": kind,
        "sample": ("first %d keyframes / %d points / %d observations of the same scene, %d LM iterations of the "
                   "reference solver (Ceres 2.0.0 SPARSE_SCHUR, %d threads); LM it/s scaled by the observation "
                   "ratio %.5f to the full workload (linear-cost assumption)" %
                   (smp.n_poses, smp.n_landmarks, smp.n_obs, lm_its, cores, scale)),
        "sample_lm_it_per_s": it_per_s, "sample_obs": int(smp.n_obs),
        "resjac_obs_per_s": resjac, "sample_wall_s": wall,
        "linear_solver_s_per_solve": s.linear_solver_time_in_seconds / max(s.num_linear_solves, 1),
        "final_cost": s.final_cost,
    }


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import pba_b200 as pb

    if a.impl == "reference":
        # the reference's own CPU path; rank 0 alone runs it
        if rank != 0:
            return
        # same scene structure as the GPU arm; only the sample's keyframes are rendered (CPU renderer)
        prob, gt = pb.make_scene(a.mode, a.kf, a.pts, a.model, render=False)
        if a.mode == 1:
            from pba_b200 import _ffi
            n_img = min(a.kf, a.sample_kf)
            _ffi.load_synth().pba_synth_render(C.byref(gt["params"]), 0, n_img, 752, _ffi.ptr(prob.images, C.c_uint8))
        full_obs = int(prob.n_obs)
        if a.warmup > 0:
            run_cpu_reference(prob, argparse.Namespace(**{**vars(a), "sample_kf": 14}), 1, full_obs)
        cb = run_cpu_reference(prob, a, max(a.steps, 1), full_obs)
        line = {
            "impl": "reference", "metric": "lm_iterations_per_s", "value": cb["value"], "unit": "LM it/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 / max(cb["value"], 1e-30), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, full_obs), "sample": cb["sample"]},
            "resjac_obs_per_s": cb["resjac_obs_per_s"],
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "LM it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic workload (identical on every rank; images ray-cast on this rank's GPU) ----
    t0 = time.time()
    prob, gt = pb.make_scene(a.mode, a.kf, a.pts, a.model, render=False)
    if a.mode == 1:
        pinned = torch.empty((a.kf, 480, 752), dtype=torch.uint8, pin_memory=True)
        prob.images = pinned.numpy()
        from pba_b200 import _ffi
        _ffi.check(_ffi.load_lib().pba_synth_render_gpu(C.byref(gt["params"]), 0, a.kf, 752,
                                                        _ffi.ptr(prob.images, C.c_uint8)), "render")
    t_scene = time.time() - t0
    hub = 9.0 if a.mode == 1 else 1.0
    opts = pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, device=local_rank, profile=True,
                                      solver=a.solver)

    # ---- resident engine ----
    t0 = time.time()
    eng = pb.Engine(prob, opts, rank=rank, world_size=world)
    t_create = time.time() - t0
    if world > 1:
        from pba_b200.engine import nccl_unique_id
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        comm_id = bytes(idt.cpu().numpy().tolist())
        eng.comm_init(comm_id)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    radius = 1e4
    clocks = ClockSampler(local_rank)
    clocks.start()  # sampling spans warm-up + timed region (all under load)
    t_w = time.time()
    n_warm = max(a.warmup, 3)
    for _ in range(n_warm):
        it = eng.lm_iterate(radius)
    # keep the GPU loaded for >= 1 s before timing so nvidia-smi gets samples; every rank must
    # run the same number of steps (each contains the all-reduce), so rank 0 decides
    extra = torch.tensor([int(max(0.0, 1.0 - (time.time() - t_w)) / max((time.time() - t_w) / n_warm, 1e-4)) + 1],
                         dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(extra, 0)
    for _ in range(int(extra.item())):
        it = eng.lm_iterate(radius)
    n_warm += int(extra.item())
    # timed region: CUDA events only around the roofline kernel (K1); bracketing all ~60 launches of a
    # step costs ~0.3 ms of gaps per step, so the per-kernel table is taken in a separate pass below
    eng.set_profile(2)
    eng.reset_kernel_stats()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(a.steps):
        it = eng.lm_iterate(radius)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    stats = eng.kernel_stats()
    # per-kernel breakdown: same step, every kernel bracketed (not part of the timed region)
    n_prof = min(a.steps, 5)
    eng.set_profile(1)
    eng.lm_iterate(radius)  # re-align the ranks (the all-reduce's event time includes waiting for the slowest)
    if world > 1:
        dist.barrier()
    eng.reset_kernel_stats()
    for _ in range(n_prof):
        eng.lm_iterate(radius)
    eng.synchronize()
    stats_all = eng.kernel_stats()
    eng.set_profile(2)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    k1ms = torch.tensor([stats.get("residual_jacobian", (0, 0.0))[1]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(k1ms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    ms_per_step = ms / a.steps
    value = a.steps / (ms * 1e-3)
    launches = int(sum(v[0] for v in stats.values()))
    k1_ms = float(k1ms.item()) / a.steps  # slowest rank's average K1 launch
    n_obs_local = eng.n_obs_local
    peak, peak_src = measured_peaks()
    bpo = K1_BYTES_PER_OBS[a.mode]
    achieved = n_obs_local * bpo / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            key = "%d_%d_%d" % (a.mode, a.kf, a.pts)
            traffic = tj.get(key)
        except Exception:
            traffic = None
    solver_used = ("bcr (block cyclic reduction)" if "bcr" in stats else "band_cholesky" if "band_cholesky" in stats
                   else "pcg" if "pcg" in stats else "dense_cholesky_dmma")

    # ---- end to end: the drop-in call with HOST buffers ----
    e2e = None
    if not a.no_e2e:
        p2 = prob.copy()
        o2 = pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, device=local_rank, solver=a.solver,
                                        max_num_iterations=20)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        if world == 1:
            s2 = pb.bundle_adjustment(p2, o2)  # pba_solve: flatten + H2D + LM + D2H
        else:
            # same id as the resident engine: the engine keeps one NCCL communicator per process and
            # id (communicator creation is process set-up, like torch.distributed's, not part of a solve)
            e2 = pb.Engine(p2, o2, rank=rank, world_size=world)
            e2.comm_init(comm_id)
            s2 = e2.minimize()
            e2.get_state()
            e2.close()
        torch.cuda.synchronize()
        t_e2e = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        lm_its = max(s2.num_iterations - 1, 1)
        h2d = int(prob.poses.nbytes + prob.inv_depth.nbytes + prob.lm_host.nbytes + prob.lm_host_uv.nbytes
                  + prob.lm_obs_ptr.nbytes + prob.obs_target.nbytes * 3 + prob.obs_target.nbytes * 4
                  + (prob.images.nbytes if prob.images is not None else 0)
                  + (prob.obs_uv.nbytes if prob.obs_uv is not None else 0)
                  + (prob.affine.nbytes if prob.affine is not None else 0))
        d2h = int(prob.poses.nbytes + prob.inv_depth.nbytes + (prob.affine.nbytes if prob.affine is not None else 0)
                  + 17 * 8 * (2 * lm_its + 2))
        e2e = {"value": lm_its / float(t_e2e.item()), "unit": "LM it/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "call": ("pba_solve(max_num_iterations=20) on host buffers" if world == 1 else
                        "pba_create + pba_comm_init (cached communicator) + pba_minimize(20) + pba_get_state on host buffers, per rank"),
               "lm_iterations": lm_its, "wall_s": float(t_e2e.item()), "setup_s": s2.setup_time_in_seconds,
               "minimizer_s": s2.minimizer_time_in_seconds, "solve_total_s": s2.total_time_in_seconds,
               "final_cost": s2.final_cost, "initial_cost": s2.initial_cost,
               "termination": {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}.get(s2.termination_type)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_baseline = run_cpu_reference(prob, a, a.cpu_iters, prob.n_obs)

    if rank == 0:
        line = {
            "metric": "lm_iterations_per_s", "value": value, "unit": "LM it/s", "n_gpus": world, "steps": a.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(a, prob.n_obs),
                "step": "one full LM iteration (pba_lm_iterate): J+r eval, Schur/RCS build, solve, back-substitution, "
                        "model cost, candidate cost; state not advanced",
                "l2": "inputs larger than L2: stored Jacobian planes %.1f GB + images %.2f GB per GPU vs 126 MB L2" %
                      (n_obs_local * 8 * 10 * 8 / 1e9 if a.mode == 1 else n_obs_local * 2 * 8 * 8 / 1e9,
                       (prob.images.nbytes if prob.images is not None else 0) / 1e9),
                "rcs_solver": solver_used, "partition": "landmarks by observation count, %d shard(s)" % world,
                "scene_s": t_scene, "create_s": t_create,
            },
            "resjac_obs_per_s": prob.n_obs / (k1_ms * 1e-3) if k1_ms > 0 else None,
            "clocks": clk, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_eval_photo<true> (residual_jacobian)" if a.mode == 1 else "k_eval_geom<true>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_obs": bpo,
                         "obs_per_launch": int(n_obs_local), "ms_per_launch": k1_ms,
                         "frac_without_schur_record": (n_obs_local * K1_BYTES_PER_OBS_JR_ONLY[a.mode] / (k1_ms * 1e-3) / 1e9 / peak
                                                       if k1_ms > 0 else 0.0),
                         "materialised_equivalent": {
                             "bytes_per_obs": K1_BYTES_PER_OBS_MATERIALISED[a.mode],
                             "gbps": (n_obs_local * K1_BYTES_PER_OBS_MATERIALISED[a.mode] / (k1_ms * 1e-3) / 1e9
                                      if k1_ms > 0 else 0.0),
                             "note": "throughput a kernel that stored the full 8x15 Jacobian would need for the same launch time"}},
            "kernels_ms_per_step": {k: v[1] / n_prof for k, v in sorted(stats_all.items(), key=lambda kv: -kv[1][1])},
            "kernels_ms_per_step_source": "%d extra steps after the timed region with CUDA events around every kernel" % n_prof,
            "last_iteration": {"cost": it["cost"], "cost_change": it["cost_change"],
                               "relative_decrease": it["relative_decrease"],
                               "linear_solver_iterations": it["linear_solver_iterations"]},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
