#!/bin/bash
# N=4 set-up breakdown (PBA_TIMING) of both end-to-end modes
mkdir -p gpurun_out
PBA_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/n4_timing.json 2> gpurun_out/n4_timing.err
echo rc=$?
grep -c pba_create gpurun_out/n4_timing.err
tail -c 1500 gpurun_out/n4_timing.json
