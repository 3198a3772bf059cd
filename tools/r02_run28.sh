#!/bin/bash
# one-launch cooperative dense Cholesky: solver tests, EuRoC / grid parity, config-1 bench lines (new vs PBA_CHOL_V1)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_euroc.py tests/test_structure.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_scale.py -q -m gpu -x -k "non_banded" 2>&1 | tail -3
for v in new v1; do
  if [ $v = v1 ]; then export PBA_CHOL_V1=1; else unset PBA_CHOL_V1; fi
  timeout 300 python bench.py --workload euroc_geom --steps 20 --warmup 5 --no-cpu-baseline > $O/euroc_geom_$v.json 2> $O/euroc_geom_$v.err; tail -1 $O/euroc_geom_$v.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/euroc_geom_$v.json') if l.startswith('{')][-1])
print('$v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k:round(x,4) for k,x in d['kernels_ms_per_step'].items() if x>0.01}, d.get('parity'))
PY
done
