#!/bin/bash
mkdir -p gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
  python -c "
import json
d=json.load(open('gpurun_out/scale_$N.json'))
print('N=$N value %.2f it/s  ms/step %.3f  e2e %.2f  K1 frac %.3f'%(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value',0),d['roofline']['frac']))
print('   ', {k:round(v,3) for k,v in list(d['kernels_ms_per_step'].items())[:9]})
"
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_1.err | tail -1 > gpurun_out/scale_1.json
python -c "
import json
d=json.load(open('gpurun_out/scale_1.json'))
print('N=1 value %.2f it/s  ms/step %.3f  e2e %.2f  K1 frac %.3f'%(d['value'],d['ms_per_step'],d.get('e2e',{}).get('value',0),d['roofline']['frac']))
"
