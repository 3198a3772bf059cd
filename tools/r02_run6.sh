#!/bin/bash
# 2 GPUs: multi-rank tests, single-process multi-GPU, N=2 bench, BCR launch list
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/box2.txt
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_multi.log 2>&1; echo "pytest rc $?" >> $O/pytest_multi.log
tail -15 $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2> $O/scale_2.err | tail -1 > $O/scale_2.json
tail -3 $O/scale_2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/scale_2.json').read().strip().splitlines()[-1])
print('N=2 value %.2f ms/step %.3f'%(d['value'],d['ms_per_step']))
print('e2e',d.get('e2e'))
print('parity',d.get('parity'))
print({k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})
PY
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
