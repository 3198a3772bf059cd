#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
PBA_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2> $O/scale_2.err | tail -1 > $O/scale_2.json
grep -E "pba_minimize|Error|error" $O/scale_2.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/scale_2.json').read().strip().splitlines()[-1])
print('N=2 value %.2f ms/step %.3f'%(d['value'],d['ms_per_step']))
e=d.get('e2e')
print('e2e', {k:e.get(k) for k in ('value','wall_s','setup_s','minimizer_s')})
print('single', e.get('single_process'))
print('parity',d.get('parity'))
PY
PBA_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/scale_1.json 2> $O/scale_1.err
grep -E "pba_minimize" $O/scale_1.err | tail -3; grep pba_create $O/scale_1.err | tail -10
python - <<'PY'
import json
d=json.loads(open('gpurun_out/scale_1.json').read().strip().splitlines()[-1])
print('N=1 value %.2f ms/step %.3f'%(d['value'],d['ms_per_step']))
e=d.get('e2e')
print('e2e', {k:e.get(k) for k in ('value','wall_s','setup_s','minimizer_s')})
print({k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})
PY
