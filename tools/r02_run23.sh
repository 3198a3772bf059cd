#!/bin/bash
# full GPU suite (with durations), smoke, default bench: state check after the container was re-created
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > $O/pytest_gpu.log 2>&1
tail -30 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench_n1.json 2> $O/bench_n1.err; tail -4 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['kernels_ms_per_step'])
print(d.get('parity'), d.get('cpu_baseline'))
PY
