#!/bin/bash
# K2 in two halves: variants 3 / 4 / 5 (5 / 6 / 8 CTAs per SM) against the default
mkdir -p gpurun_out
for v in 0 3 4 5; do
  PBA_K2_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/k2v$v.json 2> gpurun_out/k2v$v.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/k2v$v.json') if l.startswith('{')][-1])
print('K2 variant $v', round(d['value'],2), round(d['ms_per_step'],3), 'cost_only', round(d['kernels_ms_per_step']['cost_only'],4), 'final it cost', d.get('last_iteration',{}).get('cost_change'))
PY
done
