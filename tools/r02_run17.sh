#!/bin/bash
# source-level profile of k_b2_top (one CTA: Cholesky + one-tile sweeps) and one upper-level k_b2_fs
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:k_b2_top -s 4 -c 1 -o $O/b2_top -f $B > $O/ncu_b2top.log 2>&1
echo rc=$?
ls -la $O/b2_top.ncu-rep
