"""gpurun_out/ of tools/r02_profile.sh -> profiles/<tag>_* (bench lines, launch list + shares, ncu details per kernel,
K1 traffic).  Usage: python tools/r02_artifacts.py r02a"""
import collections, csv, json, os, re, subprocess, sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out, go = os.path.join(root, "profiles"), os.path.join(root, "gpurun_out")


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


for src, dst in (("bench_full.json", "bench_cfg4_n1.json"), ("bench_cfg2.json", "bench_cfg2.json"),
                 ("bench_cfg3_ds.json", "bench_cfg3_ds.json"), ("bench_cfg3_kb4.json", "bench_cfg3_kb4.json"),
                 ("bench_cfg5.json", "bench_cfg5.json"), ("bench_cfg1_euroc_geom.json", "bench_cfg1_euroc_geom.json"),
                 ("bench_cfg1_euroc_photo.json", "bench_cfg1_euroc_photo.json"), ("bench_grid.json", "bench_grid.json")):
    p = os.path.join(go, src)
    if os.path.exists(p):
        d = last_json(p)
        json.dump(d, open(os.path.join(out, "%s_%s" % (tag, dst)), "w"), indent=1)
        e = d.get("e2e") or {}
        print("%-20s %8.2f it/s %7.3f ms/step  K1 frac %.3f  e2e %.1f  parity %s" % (
            dst, d["value"], d["ms_per_step"], d["roofline"]["frac"], e.get("value", 0),
            {k: (("%.1e" % v) if isinstance(v, float) else v) for k, v in (d.get("parity") or {}).items()
             if k in ("final_cost_rel", "iterations_equal")}))
subprocess.check_call(["cp", os.path.join(go, "launches.csv"), os.path.join(out, tag + "_launches_bench_cfg4.csv")])
agg = collections.OrderedDict()
for r in csv.reader(open(os.path.join(go, "launches.csv"))):
    if len(r) <= 10 or not r[0].isdigit():
        continue
    m = re.search(r"(k_\w+(<[^>]*>)?)", r[4])
    key = m.group(1) if m else r[4][:40]
    if key.startswith("k_synth") or "render" in key or "k_build_quads" in key or "init" in key or "expand_edges" in key:
        continue
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e6
tot = sum(x[1] for x in agg.values())
with open(os.path.join(out, tag + "_launch_shares.csv"), "w") as f:
    f.write("ncu launch list (gpu__time_duration.sum, --clock-control none) of `bench.py --steps 2 --warmup 3 --no-e2e "
            "--no-cpu-baseline --no-parity`; cold-cache serialised times; shares exclude scene generation/upload kernels\n")
    f.write("kernel,launches,total_ms,share\n")
    for k, x in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.3f,%.1f%%\n" % (k, x[0], x[1], 100 * x[1] / tot))
print(open(os.path.join(out, tag + "_launch_shares.csv")).read()[230:900])

KEEP = re.compile(r"Duration|Elapsed Cycles|DRAM Throughput|Memory Throughput|Compute \(SM\) Throughput|Registers Per Thread|"
                  r"Theoretical Occupancy|Achieved Occupancy|Executed Ipc Active|No Eligible|L2 Hit Rate|Mem Busy|"
                  r"Dynamic Shared Memory|Issue Slots Busy")
traffic = {}
for name in ("k1_full", "k2_full", "gram_full", "syrk_full", "gather_full", "backsub_full", "b2_fs_l0_full", "b2_fs_l3_full",
             "b2_reduce_l0_full", "k1_geom_full", "chol_syrk_full", "chol_coop_full", "chol_coop_grid_full"):
    rep = os.path.join(go, name + ".ncu-rep")
    if os.path.exists(rep):
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    elif os.path.exists(os.path.join(go, name + ".details.txt")):  # extracted on the GPU box (r02d_profile.sh)
        det = open(os.path.join(go, name + ".details.txt")).read()
        raw = open(os.path.join(go, name + ".raw.csv")).read()
    else:
        continue
    open(os.path.join(out, "%s_%s_details.txt" % (tag, name.replace("_full", ""))), "w").write(det)
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]

    def col(metric):
        return rows[2][h.index(metric)], rows[1][h.index(metric)]
    rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tr = float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]]
    traffic[name] = tr
    extra = {}
    for metric in ("sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                   "sm__inst_executed_pipe_fp64_op_dmma.sum", "smsp__inst_executed_pipe_fp64_op_dmma.sum",
                   "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"):
        if metric in h:
            extra[metric] = col(metric)
    if name == "k1_full":
        open(os.path.join(out, tag + "_k1_eval_photo_full_raw.csv"), "w").write(raw)
    print("==", name, "DRAM traffic %.3f GB" % (tr / 1e9), extra)
    for line in det.splitlines():
        if KEEP.search(line) and "OPT" not in line and "INF" not in line:
            print("   ", line.strip()[:110])
tp = os.path.join(out, "k1_traffic.json")
tj = json.load(open(tp))
if "k1_full" in traffic:
    tj["1_2000_2000000"] = traffic["k1_full"]
if "k1_geom_full" in traffic:
    tj["0_1000_1000000"] = traffic["k1_geom_full"]
tj["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the residual/Jacobian kernel at the named "
               "workload (key = mode_keyframes_points), ncu --set full, profiles/%s_*; bench.py scales it by the launch's "
               "share of the observations at N > 1" % tag)
json.dump(tj, open(tp, "w"), indent=1)
