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
    ap.add_argument("--workload", default="synthetic", choices=["synthetic", "euroc_geom", "euroc_photo", "grid"],
                    help="euroc_*: BASELINE config 1, the map built from data/euroc_V1 by tools/euroc/ (fixtures under "
                         "tests/golden/): geometric BA of the whole map / photometric BA on its first 12 keyframes")
    ap.add_argument("--sample-kf", type=int, default=48, help="keyframes in the CPU-baseline sample")
    ap.add_argument("--cpu-iters", type=int, default=3, help="LM iterations of the cpu_baseline leg")
    ap.add_argument("--ref-max-iters", type=int, default=2,
                    help="--impl reference: LM iterations of the full-size CPU run (about a minute each at 2k x 2M)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def load_euroc(a):
    """BASELINE config 1 fixtures (tools/euroc/build_map.py); `grid`: the non-banded lawn-mower flight of
    tests/golden/scale_grid.npz (27 x 27 keyframes, 4,362 unknowns in the reduced camera system)."""
    import pba_b200 as pb
    if a.workload == "grid":
        return pb.make_grid_scene(27, 27, 30000)[0]
    g = np.load(os.path.join(ROOT, "tests", "golden", "euroc_v1_map.npz" if a.workload == "euroc_geom" else "euroc_v1_photo.npz"))
    photo = a.workload == "euroc_photo"
    return pb.Problem(int(g["mode"]), g["poses"], g["pose_fixed"], g["pose_calib"], g["calib_model"], g["intrinsics"],
                      g["inv_depth"], g["lm_host"], g["lm_host_uv"], g["lm_obs_ptr"], g["obs_target"],
                      None if photo else g["obs_uv"], g["images"] if photo else None, g["affine"] if photo else None)


def workload_name(a, n_obs):
    if a.workload == "grid":
        return ("geometric reprojection (Huber 1) BA on a lawn-mower flight, 27 x 27 keyframes x 30000 landmarks (%d residual "
                "blocks): reduced camera system NOT banded under any camera order (4,362 unknowns, half-bandwidth 113 "
                "cameras after reverse Cuthill-McKee), solved by the exact dense Cholesky" % n_obs)
    if a.workload != "synthetic":
        return ("BASELINE config 1: %s BA on the map built from the bundled data/euroc_V1 stereo keyframes (%s, %d residual "
                "blocks, double-sphere model calibrated from data/euroc_calib; tools/euroc/)" %
                ("photometric" if a.workload == "euroc_photo" else "geometric reprojection (Huber 1)",
                 "first 12 keyframes with their images" if a.workload == "euroc_photo" else "152 cameras, 3,930 landmarks", n_obs))
    kind = "photometric (8-px pattern, affine brightness, Huber 9)" if a.mode == 1 else "geometric reprojection (Huber 1)"
    return "%s BA, %d KF x %d pts (%d obs), %s 752x480, synthetic textured wall" % (kind, a.kf, a.pts, n_obs, a.model)


def arm_independent_config(a, n_obs):
    """`config` is the same dict in both arms (the driver compares them): it names the workload and what a
    step is; everything arm-specific lives under `detail`."""
    return {
        "workload": workload_name(a, n_obs),
        "step": "one Levenberg-Marquardt iteration: residual+Jacobian evaluation, Schur elimination to the reduced "
                "camera system, RCS solve, back-substitution, candidate cost",
        "l2": ("inputs larger than L2 (Jacobian %.1f GB + images %.2f GB vs 126 MB)" %
               (n_obs * (8 * 15 if a.mode == 1 else 2 * 13) * 8 / 1e9, (a.kf * 480 * 752 / 1e9) if a.mode == 1 else 0.0))
              if not small_workload(a, n_obs) else
              "working set smaller than 2 x L2: L2 flushed (256 MB written) before every timed iteration, each iteration "
              "timed by its own CUDA events",
    }


def small_workload(a, n_obs):
    return n_obs * (8 * 15 if a.mode == 1 else 2 * 13) * 8 < 2 * 126e6


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return float(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


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
    return prob.prefix_keyframes(n_kf_s)


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
    smp = sample_problem(prob, min(a.sample_kf, prob.n_poses)) if a.workload == "synthetic" else prob.copy()
    opts = of.default_options(huber_parameter=hub, max_num_iterations=iters)
    t0 = time.time()
    s = of.solve(lib, smp, opts, threads=cores)
    wall = time.time() - t0
    lm_its = max(s.num_iterations - 1, 1)
    it_per_s = lm_its / max(s.minimizer_time_in_seconds, 1e-9)
    resjac = smp.n_obs * s.num_jacobian_evaluations / max(s.jacobian_evaluation_time_in_seconds, 1e-9)
    scale = smp.n_obs / float(n_obs_full)
    return {
        "value": it_per_s * scale, "unit": "LM it/s", "cores": cores, "kind": kind, "extrapolated": True,
        "sample": ("first %d keyframes / %d points / %d observations of the same scene, %d LM iterations of the "
                   "reference solver (Ceres 2.0.0 SPARSE_SCHUR, %d threads); EXTRAPOLATED: LM it/s scaled by the "
                   "observation ratio %.5f to the full workload (linear-cost assumption) - the measured full-size "
                   "number is the `--impl reference` arm" %
                   (smp.n_poses, smp.n_landmarks, smp.n_obs, lm_its, cores, scale)),
        "sample_lm_it_per_s": it_per_s, "sample_obs": int(smp.n_obs),
        "resjac_obs_per_s": resjac, "sample_wall_s": wall,
        "linear_solver_s_per_solve": s.linear_solver_time_in_seconds / max(s.num_linear_solves, 1),
        "final_cost": s.final_cost,
    }, smp, s, opts


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import pba_b200 as pb

    if a.impl == "reference":
        # The reference's own CPU path (oracle/_ref: unmodified visnav functor + vendored Ceres 2.0.0,
        # SPARSE_SCHUR, all host threads) on the FULL workload; rank 0 alone runs it.  Every number on the line
        # is measured (Solver::Summary), nothing is extrapolated.  An LM iteration of the 2k x 2M scene costs the
        # CPU about a minute, so the run does min(--steps, --ref-max-iters) iterations and reports that count
        # as `steps` (`steps_requested` keeps the flag).
        if rank != 0:
            return
        import oracle_ffi as of
        kind = "reference" if of.have_ref() else "port"
        lib = "ref" if kind == "reference" else "oracle"
        cores = os.cpu_count() or 1
        hub = 9.0 if a.mode == 1 else 1.0
        t0 = time.time()
        if a.workload != "synthetic":
            prob, gt = load_euroc(a), None
            a.mode = prob.mode
            hub = 9.0 if a.mode == 1 else 1.0
        else:
            prob, gt = pb.make_scene(a.mode, a.kf, a.pts, a.model, render=False)
        full_obs = int(prob.n_obs)
        # Ceres holds the Jacobian (960 B / photometric block) plus ~0.75 KB of bookkeeping per block
        # (measured here: 31 GB resident at 18M blocks); if the host cannot hold that, the run uses the longest
        # keyframe prefix that fits and says so (the config then differs from the GPU arm's).
        need_gb = full_obs * ((8 * 15 * 8 + 800) if a.mode == 1 else (2 * 13 * 8 + 700)) / 1e9 + 2.0
        avail_gb = mem_available_gb()
        run_prob, note = prob, "full workload"
        n_pref, pref = of.largest_reference_prefix(prob)
        if pref is not prob:
            # Ceres 2.0.0 counts Jacobian non-zeros in an int: 17,959,243 blocks x 120 entries = 2.155e9 > 2^31 - 1
            # aborts in block_sparse_matrix.cc:80, so the reference can only run this prefix of the scene
            run_prob = pref
            note = ("vendored Ceres 2.0.0 holds at most 2^31-1 Jacobian entries (block_sparse_matrix.cc:80; the full "
                    "scene has %.3fe9): longest keyframe prefix that fits = %d of %d keyframes, %.2f %% of the "
                    "observations" % (full_obs * prob.res_per_obs * prob.cols_per_obs / 1e9, n_pref, prob.n_poses,
                                      100.0 * pref.n_obs / full_obs))
            need_gb *= pref.n_obs / float(full_obs)
        if avail_gb and need_gb > 0.9 * avail_gb:
            frac = 0.9 * avail_gb / need_gb
            run_prob = sample_problem(prob, max(8, int(run_prob.n_poses * frac)))
            note = ("host has %.0f GB available, the full problem needs ~%.0f GB in Ceres: first %d keyframes / %d "
                    "observations" % (avail_gb, need_gb, run_prob.n_poses, run_prob.n_obs))
        if a.mode == 1 and gt is not None:
            from pba_b200 import _ffi
            _ffi.load_synth().pba_synth_render(C.byref(gt["params"]), 0, run_prob.n_poses, 752,
                                               _ffi.ptr(prob.images, C.c_uint8))
        t_scene = time.time() - t0
        if a.warmup > 0:  # page the library in, spin the thread pool up
            run_cpu_reference(prob, argparse.Namespace(**{**vars(a), "sample_kf": 14}), 1, full_obs)
        iters = max(1, min(a.steps, a.ref_max_iters))
        opts = of.default_options(huber_parameter=hub, max_num_iterations=iters)
        t0 = time.time()
        sm = of.solve(lib, run_prob, opts, threads=cores)
        wall = time.time() - t0
        its = sm.iterations
        lm_its = max(len(its) - 1, 1)
        # iteration 0 is the initial evaluation; the LM iterations are what lies after it
        t_lm = its[-1]["cumulative_time_in_seconds"] - its[0]["cumulative_time_in_seconds"] if len(its) > 1 else \
            sm.minimizer_time_in_seconds
        value = lm_its / max(t_lm, 1e-9)
        resjac = run_prob.n_obs * sm.num_jacobian_evaluations / max(sm.jacobian_evaluation_time_in_seconds, 1e-9)
        cb = {"value": value, "unit": "LM it/s", "cores": cores, "kind": kind, "extrapolated": False,
              "sample": "%s: %d keyframes / %d points / %d observations, %d LM iterations (Ceres 2.0.0 SPARSE_SCHUR, "
                        "%d threads), times from Solver::Summary" %
                        (note, run_prob.n_poses, run_prob.n_landmarks, run_prob.n_obs, lm_its, cores),
              "resjac_obs_per_s": resjac, "wall_s": wall, "scene_s": t_scene,
              "problem_build_and_preprocess_s": wall - sm.minimizer_time_in_seconds,
              "minimizer_s": sm.minimizer_time_in_seconds,
              "jacobian_s_per_eval": sm.jacobian_evaluation_time_in_seconds / max(sm.num_jacobian_evaluations, 1),
              "residual_s_per_eval": sm.residual_evaluation_time_in_seconds / max(sm.num_residual_evaluations, 1),
              "linear_solver_s_per_solve": sm.linear_solver_time_in_seconds / max(sm.num_linear_solves, 1),
              "iteration_s": [it["iteration_time_in_seconds"] for it in its],
              "iteration_cost": [it["cost"] for it in its],
              "initial_cost": sm.initial_cost, "final_cost": sm.final_cost}
        line = {
            "impl": "reference", "metric": "lm_iterations_per_s", "value": value, "unit": "LM it/s",
            "n_gpus": a.gpus, "steps": lm_its, "steps_requested": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 / max(value, 1e-30), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # `config` names the benchmark workload (identical in both arms); what the reference could actually
            # hold of it is in `reference_workload` (and in cpu_baseline.sample)
            "config": arm_independent_config(a, full_obs),
            "reference_workload": {"note": note, "keyframes": int(run_prob.n_poses), "points": int(run_prob.n_landmarks),
                                   "observations": int(run_prob.n_obs),
                                   "fraction_of_observations": run_prob.n_obs / float(full_obs)},
            "resjac_obs_per_s": resjac,
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "LM it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "with_problem_build": lm_its / wall,
                    "note": "value = the line's own; with_problem_build = LM iterations / wall clock of ceres::Problem "
                            "build + Solve on host containers"},
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
    if a.workload != "synthetic":
        prob, gt = load_euroc(a), None
        a.mode = prob.mode
    else:
        prob, gt = pb.make_scene(a.mode, a.kf, a.pts, a.model, render=False)
    if a.mode == 1 and gt is not None:
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
    if small_workload(a, prob.n_obs):
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")  # 256 MB > 126 MB of L2
        ms = 0.0
        for _ in range(a.steps):
            flush.zero_()
            torch.cuda.synchronize()
            ev0.record(stream)
            it = eng.lm_iterate(radius)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms += ev0.elapsed_time(ev1)
        del flush
    else:
        ev0.record(stream)
        for _ in range(a.steps):
            it = eng.lm_iterate(radius)
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
    if world > 1:
        dist.barrier()
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
            if traffic is not None and prob.n_obs > 0:
                # the capture is one launch over ALL observations (N = 1); a shard's launch moves its share
                traffic = traffic * n_obs_local / float(prob.n_obs)
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
        # N = 1: the drop-in call, pba_solve.  N > 1 (one process per GPU under torchrun): every rank makes the
        # split calls on its landmark shard (pba_create + pba_comm_init + pba_minimize + pba_get_state) -> e2e.value,
        # max over ranks; and, separately, rank 0 ALONE drives all N GPUs through the drop-in call the reference's
        # single-threaded caller would make (src/sfm.cpp:1903-1913), pba_solve(num_gpus = N) -> e2e.single_process.
        # Steady state, as in the SfM loop's repeated optimize() calls (src/sfm.cpp:1153-1155): an untimed
        # one-iteration solve first wherever the pools (device arena, pinned staging) have not seen these sizes.
        def run_per_rank(opts):
            tt = [time.time()]
            e2 = pb.Engine(p2, opts, rank=rank, world_size=world)
            tt.append(time.time())
            e2.comm_init(comm_id)  # same id as the resident engine: the communicator is cached per process
            tt.append(time.time())
            sm = e2.minimize()
            tt.append(time.time())
            p2.poses[:] = e2.get_state()[0]
            tt.append(time.time())
            e2.close()
            tt.append(time.time())
            if os.environ.get("PBA_TIMING"):
                print("[bench e2e rank %d] create %.1f comm_init %.1f minimize %.1f get_state %.1f close %.1f ms" %
                      ((rank,) + tuple(1e3 * (tt[i + 1] - tt[i]) for i in range(5))), file=sys.stderr)
            return sm
        if world == 1:
            pb.bundle_adjustment(prob.copy(), pb.BundleAdjustmentOptions(
                verbosity_level=0, huber_parameter=hub, device=local_rank, solver=a.solver, max_num_iterations=1))
        else:
            p_keep = p2
            p2 = prob.copy()
            run_per_rank(pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, device=local_rank,
                                                    solver=a.solver, max_num_iterations=1))
            p2 = p_keep
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        s2 = pb.bundle_adjustment(p2, o2) if world == 1 else run_per_rank(o2)
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
                        "pba_create + pba_comm_init (cached communicator) + pba_minimize(20) + pba_get_state on host "
                        "buffers, every rank on its shard (max over ranks)"),
               "bytes_are_per": "solve: one call = set-up + %d LM iterations; the copies happen once per call, "
                                "not once per iteration%s" % (lm_its, "" if world == 1 else "; per rank: its shard + all keyframes"),
               "excludes": None if world == 1 else "NCCL communicator creation (cached per process: pba_multi_gpu_init / pba_comm_init)",
               "state": "steady: device arena and pinned staging pools warm (an earlier solve / the resident engine had "
                        "the same sizes), as in the SfM loop's repeated optimize() calls",
               "lm_iterations": lm_its, "wall_s": float(t_e2e.item()), "setup_s": s2.setup_time_in_seconds,
               "minimizer_s": s2.minimizer_time_in_seconds, "solve_total_s": s2.total_time_in_seconds,
               "final_cost": s2.final_cost, "initial_cost": s2.initial_cost,
               "termination": {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}.get(s2.termination_type)}
        if world > 1 and pb.device_count() >= world:
            # the single-process drop-in call; the other ranks wait on the rendezvous store (no GPU kernel, no spinning)
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                pb.multi_gpu_init(0, world)  # NCCL communicators of this path: process set-up, cached
                o3 = pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, device=0, solver=a.solver,
                                                max_num_iterations=1, num_gpus=world)
                for _ in range(2):  # the pinned staging pool reaches its concurrent capacity on the second call
                    pb.bundle_adjustment(prob.copy(), o3)
                o3.max_num_iterations = 20
                p3 = prob.copy()
                t0 = time.time()
                s3 = pb.bundle_adjustment(p3, o3)
                t_sp = time.time() - t0
                its3 = max(s3.num_iterations - 1, 1)
                e2e["single_process"] = {
                    "value": its3 / t_sp, "unit": "LM it/s", "wall_s": t_sp, "lm_iterations": its3,
                    "call": "pba_solve(max_num_iterations=20, num_gpus=%d) on host buffers from ONE process" % world,
                    "setup_s": s3.setup_time_in_seconds, "minimizer_s": s3.minimizer_time_in_seconds,
                    "final_cost": s3.final_cost, "poses_abs_max_vs_per_rank": float(np.abs(p3.poses - p2.poses).max())}
                store.set("pba_single_process_done", "1")
            else:
                store.wait(["pba_single_process_done"])

    # ---- CPU baseline (bounded sample) and parity of the GPU path against it on the SAME sample ----
    cpu_baseline, parity = None, {}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_baseline, smp_cpu, s_cpu, _ = run_cpu_reference(prob, a, a.cpu_iters, prob.n_obs)
        if not a.no_parity:
            smp_gpu = sample_problem(prob, min(a.sample_kf, prob.n_poses)) if a.workload == "synthetic" else prob.copy()
            s_gpu = pb.bundle_adjustment(smp_gpu, pb.BundleAdjustmentOptions(
                verbosity_level=0, huber_parameter=hub, device=local_rank, solver=a.solver,
                max_num_iterations=a.cpu_iters))
            ci, gi = s_cpu.iterations, s_gpu.iterations
            parity = {
                "against": "%s on the cpu_baseline sample (%d observations, %d LM iterations, same image bytes)" %
                           (cpu_baseline["kind"], smp_cpu.n_obs, a.cpu_iters),
                "final_cost_rel": abs(s_gpu.final_cost - s_cpu.final_cost) / s_cpu.final_cost,
                "iterations_equal": len(ci) == len(gi) and
                                    [i["step_is_successful"] for i in ci] == [i["step_is_successful"] for i in gi],
                "iteration_cost_rel_max": (max(abs(x["cost"] - y["cost"]) / y["cost"] for x, y in zip(gi, ci))
                                           if len(ci) == len(gi) else None),
                "poses_abs_max": float(np.abs(smp_gpu.poses - smp_cpu.poses).max()),
                "inv_depth_abs_max": float(np.abs(smp_gpu.inv_depth - smp_cpu.inv_depth).max()),
                "gpu_final_cost": s_gpu.final_cost, "cpu_final_cost": s_cpu.final_cost,
            }
    if world > 1 and e2e is not None and not a.no_parity:
        # the sharded solve against ONE GPU solving the whole problem (rank 0; the others wait)
        if rank == 0:
            p1 = prob.copy()
            s1 = pb.bundle_adjustment(p1, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub,
                                                                     device=local_rank, solver=a.solver,
                                                                     max_num_iterations=20))
            parity = {"sharded_vs_single_rel": abs(e2e["final_cost"] - s1.final_cost) / s1.final_cost,
                      "sharded_final_cost": e2e["final_cost"], "single_gpu_final_cost": s1.final_cost,
                      "iterations_equal": s1.num_iterations - 1 == e2e["lm_iterations"],
                      "poses_abs_max": float(np.abs(p1.poses - p2.poses).max())}
        dist.barrier()

    if rank == 0:
        line = {
            "metric": "lm_iterations_per_s", "value": value, "unit": "LM it/s", "n_gpus": world, "steps": a.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": arm_independent_config(a, prob.n_obs),
            "detail": {
                "step": "pba_lm_iterate: J+r eval (K1), Schur/RCS build, solve, back-substitution, model cost, "
                        "candidate cost (K2); state not advanced, so every step does identical work",
                "rcs_solver": solver_used, "partition": "landmarks by observation count, %d shard(s)" % world,
                "collective": {"none": "none (one rank)", "nccl": "ncclAllReduce of the partial RCS + scalars",
                               "peer": "own NVLink kernels over peer-mapped buffers (csrc/peer.cu): reduce-scatter + "
                                       "push all-gather of the partial RCS, scalar table; NCCL only for set-up"}[eng.collective],
                "scene_s": t_scene, "create_s": t_create,
            },
            "resjac_obs_per_s": prob.n_obs / (k1_ms * 1e-3) if k1_ms > 0 else None,
            "clocks": clk, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_eval_photo<true> (residual_jacobian)" if a.mode == 1 else "k_eval_geom<true>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "ncu --set full capture at N=1 (profiles/k1_traffic.json) x "
                                                                "this launch's share of the observations",
                         "peak_source": peak_src, "bytes_per_obs": bpo,
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
        if parity:
            line["parity"] = parity
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
