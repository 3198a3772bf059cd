#!/bin/bash
# usage: tools/ncu_kernel.sh <kernel-regex> <out-name> [kf] [pts]
mkdir -p gpurun_out
KF=${3:-400}; PTS=${4:-400000}
CMD="python tools/probe.py --kf $KF --pts $PTS --iters 2"
$CMD > gpurun_out/plain_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s 1 -c 1 -f -o gpurun_out/$2 $CMD > gpurun_out/ncu_$2.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_$2.log
