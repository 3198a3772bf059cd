#!/bin/bash
# round 2, GPU run 1: box facts, the GPU test-suite, K1/K2/BCR A/B runs, default bench, reference arm (1 iteration)
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g; } > $O/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -40 $O/pytest_gpu.log
# memcheck of the new solver on small problems
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "reduced_camera_system or other_solvers" -p no:cacheprovider > $O/memcheck_bcr2.log 2>&1; echo "memcheck rc $?" >> $O/memcheck_bcr2.log
tail -5 $O/memcheck_bcr2.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for v in 0 1 2 3; do PBA_K1_VARIANT=$v timeout 300 $B > $O/k1_var$v.json 2> $O/k1_var$v.err; done
PBA_K1_TABLE=1 timeout 300 $B > $O/k1_table.json 2> $O/k1_table.err
for v in 1 2; do PBA_K2_VARIANT=$v timeout 300 $B > $O/k2_var$v.json 2> $O/k2_var$v.err; done
PBA_BCR_V1=1 timeout 300 $B > $O/bcr_v1.json 2> $O/bcr_v1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k1_*.json')+glob.glob('gpurun_out/k2_*.json')+glob.glob('gpurun_out/bcr_v1.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels_ms_per_step']
        print(f, 'ms/step %.3f'%d['ms_per_step'], 'K1 %.3f'%k.get('residual_jacobian',0), 'K2 %.3f'%k.get('cost_only',0), 'bcr %.3f'%k.get('bcr',0), 'frac %.3f'%d['roofline']['frac'], 'cost', d['last_iteration']['cost'], d['last_iteration']['cost_change'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err; cat $O/bench.json
PBA_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_timing.json 2> $O/timing_e2e.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 --ref-max-iters 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -2 $O/bench_ref.err; cat $O/bench_ref.json
