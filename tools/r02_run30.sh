#!/bin/bash
# cooperative Cholesky (final): solver tests, EuRoC / grid parity, config-1 bench lines (both workloads)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_euroc.py tests/test_structure.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_scale.py -q -m gpu -x -k "non_banded" --durations=3 2>&1 | tail -8
for w in euroc_geom euroc_photo; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 > $O/r02d_bench_cfg1_$w.json 2> $O/$w.err; tail -1 $O/$w.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02d_bench_cfg1_$w.json') if l.startswith('{')][-1])
print('$w', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k:round(x,4) for k,x in d['kernels_ms_per_step'].items() if x>0.01}, d.get('parity'), d.get('cpu_baseline',{}).get('value'))
PY
done
