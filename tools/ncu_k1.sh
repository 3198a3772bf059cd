#!/bin/bash
# quick K1 capture at a mid-size workload (development)
mkdir -p gpurun_out
CMD="python tools/probe.py --kf 400 --pts 400000 --iters 2"
$CMD > gpurun_out/plain_k1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_eval_photo -s 2 -c 2 -f -o gpurun_out/k1_dev $CMD > gpurun_out/ncu_k1.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_k1.log
