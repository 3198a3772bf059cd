#!/bin/bash
# phase stamps of the cooperative Cholesky (timing build: make lib EXTRA=-DPBA_CHOL_TIMING)
mkdir -p gpurun_out
timeout 300 python bench.py --workload euroc_geom --steps 30 --warmup 15 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/chol_t.json 2> gpurun_out/chol_t.err; grep "\[chol\]" gpurun_out/chol_t.err
