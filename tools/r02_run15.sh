#!/bin/bash
# BCR2 Cholesky with a chain warp: solver tests + bench kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_structure.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/b15.json 2> gpurun_out/b15.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b15.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
print({k:round(v,4) for k,v in d['kernels_ms_per_step'].items()})
PY
