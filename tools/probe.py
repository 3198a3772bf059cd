"""Development probe: per-kernel timing of one LM iteration at a given problem size."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pba_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--kf", type=int, default=50)
ap.add_argument("--pts", type=int, default=20000)
ap.add_argument("--model", default="pinhole")
ap.add_argument("--mode", type=int, default=1)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--solver", type=int, default=0)
ap.add_argument("--minimize", action="store_true")
a = ap.parse_args()
t = time.time()
prob, gt = pb.make_scene(a.mode, a.kf, a.pts, a.model, gpu_render=True)
print("scene: %.2fs  obs=%d" % (time.time() - t, prob.n_obs), flush=True)
hub = 9.0 if a.mode == 1 else 1.0
t = time.time()
eng = pb.Engine(prob, pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, profile=True, solver=a.solver))
print("create: %.2fs" % (time.time() - t), flush=True)
eng.lm_iterate(1e4)
eng.reset_kernel_stats()
t = time.time()
for i in range(a.iters):
    it = eng.lm_iterate(1e4)
eng.synchronize()
dt = (time.time() - t) / a.iters
print("lm_iterate: %.3f ms/iter  cost %.6e -> change %.4e rel_dec %.3f ls_iters %d" % (dt * 1e3, it["cost"], it["cost_change"], it["relative_decrease"], it["linear_solver_iterations"]))
st = eng.kernel_stats()
tot = sum(v[1] for v in st.values())
for k, (n, ms) in sorted(st.items(), key=lambda kv: -kv[1][1]):
    print("  %-20s launches/iter %6.1f  ms/iter %9.4f  %5.1f%%" % (k, n / a.iters, ms / a.iters, 100 * ms / max(tot, 1e-9)))
print("  sum kernels %.3f ms/iter" % (tot / a.iters))
k1 = st["residual_jacobian"][1] / a.iters * 1e-3
bytes_obs = 1088 if a.mode == 1 else 256
print("  K1: %.3e obs/s, %.1f GB/s algorithmic" % (prob.n_obs / k1, prob.n_obs * bytes_obs / k1 / 1e9))
if a.minimize:
    t = time.time()
    s = eng.minimize()
    print("minimize: %.3fs %s  its=%d  launches=%d" % (time.time() - t, s.brief_report(), s.num_iterations, s.gpu_kernel_launches))
