#!/bin/bash
# Round-2 evidence run (one GPU): GPU test-suite, bench lines for BASELINE configs 4 (default), 2, 3 (DS, KB4), 5,
# ncu launch list of the default bench, ncu --set full captures of the kernels DESIGN.md quotes.
# Per B200_PROFILING.md every profiled command first exits 0 without ncu.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
timeout 300 python bench.py --kf 50 --pts 20000 --model pinhole --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --kf 200 --pts 100000 --model ds --steps 20 --warmup 5 > $O/bench_cfg3_ds.json 2> $O/bench_cfg3_ds.err; echo "cfg3 ds rc=$?"
timeout 300 python bench.py --kf 200 --pts 100000 --model kb4 --steps 20 --warmup 5 > $O/bench_cfg3_kb4.json 2> $O/bench_cfg3_kb4.err; echo "cfg3 kb4 rc=$?"
timeout 600 python bench.py --mode 0 --kf 1000 --pts 1000000 --steps 20 --warmup 5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {  # cap <regex> <skip> <name> [extra bench args]
  local R=$1 S=$2 N=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$R -s $S -c 1 -f -o $O/$N $CMD "$@" > $O/ncu_$N.log 2>&1
  echo "capture $N rc=$?"
}
cap k_eval_photo 6 k1_full
cap k_eval_photo 7 k2_full
cap k_edge_gram 3 gram_full
cap k_schur_syrk 3 syrk_full
cap k_lm_gather 3 gather_full
cap k_backsub 3 backsub_full
cap k_b2_fs 8 b2_fs_l0_full
cap k_b2_fs 11 b2_fs_l3_full
cap k_b2_reduce 8 b2_reduce_l0_full
cap k_eval_geom 6 k1_geom_full --mode 0 --kf 1000 --pts 1000000
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --kf 400 --pts 200000 --solver 1 > $O/plain_chol.log 2>&1 &&
cap k_chol_syrk 40 chol_syrk_full --kf 400 --pts 200000 --solver 1
ls -la $O | grep ncu-rep
