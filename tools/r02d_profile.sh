#!/bin/bash
# Round-2 final evidence run (one GPU): GPU test-suite, bench lines for BASELINE configs 4 (default), 1, 2, 3 (DS, KB4), 5
# and the non-banded flight, ncu launch list of the default bench, ncu --set full captures of the kernels DESIGN.md quotes.
# Per B200_PROFILING.md every profiled command first exits 0 without ncu.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
timeout 300 python bench.py --kf 50 --pts 20000 --model pinhole --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --kf 200 --pts 100000 --model ds --steps 20 --warmup 5 > $O/bench_cfg3_ds.json 2> $O/bench_cfg3_ds.err; echo "cfg3 ds rc=$?"
timeout 300 python bench.py --kf 200 --pts 100000 --model kb4 --steps 20 --warmup 5 > $O/bench_cfg3_kb4.json 2> $O/bench_cfg3_kb4.err; echo "cfg3 kb4 rc=$?"
timeout 600 python bench.py --mode 0 --kf 1000 --pts 1000000 --steps 20 --warmup 5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"
timeout 300 python bench.py --workload euroc_geom --steps 20 --warmup 5 > $O/bench_cfg1_euroc_geom.json 2> $O/bench_cfg1_geom.err; echo "cfg1 geom rc=$?"
timeout 300 python bench.py --workload euroc_photo --steps 20 --warmup 5 > $O/bench_cfg1_euroc_photo.json 2> $O/bench_cfg1_photo.err; echo "cfg1 photo rc=$?"
timeout 300 python bench.py --workload grid --steps 20 --warmup 5 > $O/bench_grid.json 2> $O/bench_grid.err; echo "grid rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {  # cap <regex> <skip> <name> [extra bench args]: capture, extract details + raw metrics ON THE BOX, drop the report
  local R=$1 S=$2 N=$3; shift 3       # (gpurun_out/ comes back only below 64 MiB; a report is 4-18 MB)
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$R -s $S -c 1 -f -o $O/$N $CMD "$@" > $O/ncu_$N.log 2>&1
  echo "capture $N rc=$?"
  ncu -i $O/$N.ncu-rep --page details > $O/$N.details.txt 2>/dev/null
  ncu -i $O/$N.ncu-rep --page raw --csv > $O/$N.raw.csv 2>/dev/null
  rm -f $O/$N.ncu-rep
}
cap k_eval_photo 6 k1_full
cap k_backsub 3 backsub_full
cap k_b2_fs 8 b2_fs_l0_full
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --kf 400 --pts 200000 --solver 1 > $O/plain_chol.log 2>&1 &&
cap k_chol_coop 3 chol_coop_full --kf 400 --pts 200000 --solver 1
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --workload grid > $O/plain_grid.log 2>&1 &&
cap k_chol_coop 3 chol_coop_grid_full --workload grid
rm -f $O/*.ncu-rep
du -sh $O
