#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
for w in euroc_geom euroc_photo; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; tail -2 $O/bench_$w.err
  timeout 300 python bench.py --impl reference --workload $w --steps 20 --warmup 5 --ref-max-iters 20 > $O/bench_ref_$w.json 2> $O/bench_ref_$w.err; echo "ref $w rc=$?"; tail -2 $O/bench_ref_$w.err
done
timeout 300 python bench.py --kf 50 --pts 20000 --model pinhole --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 600 python bench.py --mode 0 --kf 1000 --pts 1000000 --steps 20 --warmup 5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
for n in ('euroc_geom','euroc_photo','ref_euroc_geom','ref_euroc_photo','cfg2','cfg5'):
    try:
        d=json.loads(open('gpurun_out/bench_%s.json'%n).read().strip().splitlines()[-1])
        e=d.get('e2e') or {}
        print(n, 'value %.3f %s ms/step %.4f'%(d['value'],d['unit'],d['ms_per_step']), 'steps',d['steps'],'e2e',e.get('value'), 'parity', d.get('parity'), 'solver', (d.get('detail') or {}).get('rcs_solver'), 'K1 frac', (d.get('roofline') or {}).get('frac'))
    except Exception as ex: print(n,'FAILED',ex)
PY
