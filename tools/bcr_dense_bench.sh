#!/bin/bash
# usage: tools/bcr_dense_bench.sh [-DBCR_DENSE_PROF]   (build here, run under gpurun)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iphotometric-bundle-adjustment_b200/csrc -Iinclude "$@" tools/bcr_dense_bench.cu -o tools/bcr_dense_bench 2>&1 | grep -E "error" -A4
