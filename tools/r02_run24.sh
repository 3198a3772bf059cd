#!/bin/bash
# k_backsub with prefetched landmark descriptors: solver / LM parity tests + bench (kernel table)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_structure.py tests/test_euroc.py -q -m gpu -x 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-parity"
timeout 300 $B > $O/b24.json 2> $O/b24.err; tail -2 $O/b24.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b24.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], {k:round(v,4) for k,v in d['kernels_ms_per_step'].items()})
PY
