#!/bin/bash
# BCR2 phase stamps (build with make EXTRA=-DPBA_B2_TIMING) + launch list of the k_b2 kernels
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 300 $B > $O/b16.json 2> $O/b16.err; grep "\[b2\]" $O/b16.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_b2 -s 60 -c 30 --csv --log-file $O/bcr2_launches.csv $B > $O/ncu_bcr2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/bcr2_launches.csv')) if len(r)>10 and r[0].isdigit()]
tot=0
for r in rows[:30]:
    print(r[4][:40].ljust(40), r[7], r[8], r[-1]); 
PY
