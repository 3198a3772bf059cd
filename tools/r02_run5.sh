#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_structure.py -m gpu -q --tb=short -p no:cacheprovider -k "reduced_camera or other_solvers or config3 or config5 or shuffled" > $O/pytest_bcr.log 2>&1; echo "pytest rc $?" >> $O/pytest_bcr.log
tail -8 $O/pytest_bcr.log
timeout 600 python bench.py --steps 5 --warmup 5 --no-cpu-baseline --no-parity --no-e2e > $O/bench_short.json 2> $O/bench_short.err; grep "b2" $O/bench_short.err
