#!/bin/bash
# solver tests + BCR2 phase stamps + launch list (timing build)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_structure.py -q -m gpu -x 2>&1 | tail -3
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 300 $B > $O/b16.json 2> $O/b16.err; grep "\[b2\]" $O/b16.err
python -c "
import json
d=json.loads(open('gpurun_out/b16.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'bcr', d['kernels_ms_per_step']['bcr'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_b2 -s 66 -c 22 --csv --log-file $O/bcr2_launches.csv $B > $O/ncu_bcr2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/bcr2_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:20]:
    print(r[4][:40].ljust(40), r[7], r[8], r[-1]); 
PY
