#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/probe.py --kf ${1:-2000} --pts ${2:-200000} --iters 1"
$CMD > gpurun_out/plain_list.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/list.csv $CMD > gpurun_out/ncu_list2.log 2>&1
echo rc=$?
