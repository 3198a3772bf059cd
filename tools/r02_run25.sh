#!/bin/bash
# front-end kernels: GPU tests + throughput line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frontend.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python tools/frontend_bench.py > gpurun_out/frontend_bench.json 2> gpurun_out/frontend_bench.err; tail -3 gpurun_out/frontend_bench.err; cat gpurun_out/frontend_bench.json
