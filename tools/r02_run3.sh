#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
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
PBA_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; grep pba_create $O/bench.err | tail -14; tail -2 $O/bench.err | grep -v pba_create; cat $O/bench.json
