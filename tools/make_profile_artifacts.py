"""Turn gpurun_out/{bench_full.json, launches.csv, k1_full.ncu-rep} (tools/profile.sh) into the
profiles/<tag>_* artifacts: bench line, launch list + shares, K1 ncu details/raw pages, k1_traffic.json."""
import collections, csv, json, os, re, subprocess, sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
go = os.path.join(root, "gpurun_out")
subprocess.check_call(["cp", os.path.join(go, "bench_full.json"), os.path.join(out, tag + "_bench_cfg4_n1.json")])
subprocess.check_call(["cp", os.path.join(go, "launches.csv"), os.path.join(out, tag + "_launches_bench_cfg4.csv")])
rep = os.path.join(go, "k1_full.ncu-rep")
open(os.path.join(out, tag + "_k1_eval_photo_details.txt"), "w").write(
    subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(out, tag + "_k1_eval_photo_full_raw.csv"), "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
rd, wr = float(rows[2][h.index("dram__bytes_read.sum")]), float(rows[2][h.index("dram__bytes_write.sum")])
assert rows[1][h.index("dram__bytes_read.sum")] == "Gbyte"
tp = os.path.join(out, "k1_traffic.json")
tj = json.load(open(tp))
tj["1_2000_2000000"] = (rd + wr) * 1e9
tj["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one k_eval_photo<true,pinhole> launch at the bench workload "
               "(17,959,243 obs): %.3f GB read + %.3f GB written; ncu --set full, profiles/%s_k1_eval_photo_full_raw.csv" % (rd, wr, tag))
json.dump(tj, open(tp, "w"), indent=1)
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
            "--no-cpu-baseline`; cold-cache serialised times; shares exclude scene generation/upload kernels\n")
    f.write("kernel,launches,total_ms,share\n")
    for k, x in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.3f,%.1f%%\n" % (k, x[0], x[1], 100 * x[1] / tot))
d = json.load(open(os.path.join(go, "bench_full.json")))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["wall_s"], "roofline", d["roofline"]["frac"],
      d["roofline"]["ms_per_launch"], "traffic GB", rd + wr, "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
print({a: round(b, 3) for a, b in d["kernels_ms_per_step"].items()})
print(open(os.path.join(out, tag + "_launch_shares.csv")).read()[230:560])
