#!/bin/bash
# usage: r02_scale.sh N [N ...]   (one torchrun bench per N, summaries printed)
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
  fi
  tail -2 gpurun_out/scale_$N.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1])
print('N=$N value %.2f it/s  ms/step %.3f  K1 frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))
e=d.get('e2e') or {}
print('   e2e', {k:e.get(k) for k in ('value','wall_s','setup_s','minimizer_s','call')})
print('   parity', d.get('parity'))
print('   ', {k:round(v,3) for k,v in list(d['kernels_ms_per_step'].items())[:12]})
PY
done
