#!/bin/bash
# peer-memory collectives: 2-GPU tests + N=2 bench with and without them
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_rank.py -q -m gpu -x 2>&1 | tail -5
for mode in peer nccl; do
  if [ $mode = nccl ]; then export PBA_NO_PEER=1; else unset PBA_NO_PEER; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.err
  echo "$mode rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/n2_$mode.json').read().strip().splitlines()[-1])
k=d['kernels_ms_per_step']
print('$mode', round(d['value'],2), round(d['ms_per_step'],3), 'copy', round(k.get('copy',0),4), 'e2e', round(d['e2e']['value'],1), 'sp', round(d['e2e'].get('single_process',{}).get('value',0),1), d.get('parity'))
PY
done
