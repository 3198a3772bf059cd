#!/bin/bash
# 8-GPU box: the 2-GPU multi-rank tests, then the N=8 and N=4 bench lines (peer-memory collectives)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_rank.py -q -m gpu -x 2>&1 | tail -3
bash tools/r02_scale.sh 8 4
