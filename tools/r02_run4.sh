#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_structure.py -m gpu -q --tb=short -p no:cacheprovider -k "reduced_camera or other_solvers or config3 or config4 or config5 or shuffled or golden" > $O/pytest_bcr.log 2>&1; echo "pytest rc $?" >> $O/pytest_bcr.log
tail -8 $O/pytest_bcr.log
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_b2 -c 44 --csv --log-file $O/bcr2_launches.csv $B > $O/ncu_bcr2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/bcr2_launches.csv')) if len(r)>10 and r[0].isdigit()]
tot=0
for r in rows[:22]:
    print(r[4][:52].ljust(52), r[7], r[8], r[-1]); tot+=float(r[-1])
print('sum of first solve (ns):', tot)
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_short.json 2> $O/bench_short.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_short.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'setup', d['e2e']['setup_s'], 'min', d['e2e']['minimizer_s'])
print({k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})
PY
