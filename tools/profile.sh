#!/bin/bash
# ncu evidence for the bench workload (one GPU).  Per B200_PROFILING.md: the plain
# command must exit 0 first; then the launch list; then one --set full capture of K1.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_full.err
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_eval_photo -s 6 -c 1 -f -o gpurun_out/k1_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | head -20
