#!/bin/bash
# usage: r02_run22.sh N : bench at N GPUs with the peer-memory collectives and with NCCL (PBA_NO_PEER=1)
N=$1
mkdir -p gpurun_out
for mode in peer nccl; do
  if [ $mode = nccl ]; then export PBA_NO_PEER=1; else unset PBA_NO_PEER; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/n${N}_$mode.json 2> gpurun_out/n${N}_$mode.err
  echo "$mode rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/n${N}_$mode.json').read().strip().splitlines()[-1])
k=d['kernels_ms_per_step']
print('$mode', d['detail'].get('collective','')[:20], round(d['value'],2), round(d['ms_per_step'],3), 'copy', round(k.get('copy',0),4), 'bcr', round(k.get('bcr',0),4), 'e2e', round(d['e2e']['value'],1), 'sp', round(d['e2e'].get('single_process',{}).get('value',0),1), d.get('parity',{}).get('sharded_vs_single_rel'))
PY
done
