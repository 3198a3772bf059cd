#!/bin/bash
# N=1 set-up breakdown (PBA_TIMING) + structure tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_structure.py tests/test_abi.py -q -m gpu 2>&1 | tail -3
PBA_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/n1_timing.json 2> gpurun_out/n1_timing.err
echo rc=$?
grep -E "pba_create|analyze_cameras|iterations, wall" gpurun_out/n1_timing.err | tail -40
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n1_timing.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k:d['e2e'][k] for k in ('value','wall_s','setup_s','minimizer_s')})
PY
