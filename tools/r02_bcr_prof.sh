#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_b2 -c 60 --csv --log-file $O/bcr2_launches.csv $B > $O/ncu_bcr2.log 2>&1
PBA_BCR_V1=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_bcr -c 80 --csv --log-file $O/bcr1_launches.csv $B > $O/ncu_bcr1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:k_b2_fs -s 8 -c 1 -o $O/b2_fs_l0 -f $B > $O/ncu_b2fs.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:k_b2_reduce -s 24 -c 1 -o $O/b2_reduce_l0 -f $B > $O/ncu_b2red.log 2>&1
python - <<'PY'
import csv
for f in ['gpurun_out/bcr2_launches.csv','gpurun_out/bcr1_launches.csv']:
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    print(f, len(rows))
    for r in rows[:40]:
        print(r[4][:60].ljust(60), r[7], r[8], r[-1])
PY
