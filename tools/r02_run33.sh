#!/bin/bash
# final tree: complete GPU suite + smoke + default bench line
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu_final.log 2>&1; tail -5 $O/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], d['clocks'])
PY
