#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_triangulate.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "triangulat or add_new or residual_jacobian_parity or golden" > $O/pytest_tri.log 2>&1; echo "pytest rc $?" >> $O/pytest_tri.log
tail -6 $O/pytest_tri.log
B="python bench.py --mode 0 --kf 1000 --pts 1000000 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-parity"
for v in 0 1 2 3; do PBA_K1G_VARIANT=$v timeout 300 $B > $O/k1g_var$v.json 2> $O/k1g_var$v.err; done
python - <<'PY'
import json
for v in range(4):
    try:
        d=json.loads(open('gpurun_out/k1g_var%d.json'%v).read().strip().splitlines()[-1])
        print('geom variant',v,'ms/step %.3f'%d['ms_per_step'],'K1 %.4f ms frac %.3f'%(d['roofline']['ms_per_launch'],d['roofline']['frac']),'K2 %.3f'%d['kernels_ms_per_step'].get('cost_only',0), d['clocks']['sm_mhz'])
    except Exception as e: print(v,'failed',e)
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --kf 400 --pts 200000 --solver 1"
$CMD > $O/plain_chol.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_chol_syrk -s 49 -c 1 -f -o $O/chol_syrk_first $CMD > $O/ncu_chol_first.log 2>&1; echo "chol capture rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/plain_chol.log').read().strip().splitlines()[-1])
print('dense cholesky config: ms/step', d['ms_per_step'], {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if 'chol' in k or 'dense' in k})
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/bench_quick.json 2> $O/bench_quick.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('cfg4 ms/step', d['ms_per_step'], 'K1', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['setup_s'])
PY
