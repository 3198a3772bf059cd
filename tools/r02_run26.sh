#!/bin/bash
# front-end: GPU tests, throughput line, one ncu --set full capture of the matcher (small case)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_frontend.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python tools/frontend_bench.py > $O/frontend_bench.json 2> $O/frontend_bench.err; tail -3 $O/frontend_bench.err; cat $O/frontend_bench.json
CMD="python tools/frontend_bench.py --images 48 --ref-pairs 0 --repeat 1"
timeout 300 $CMD > $O/fe_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_match_best -s 1 -c 1 -o $O/r02d_match_best $CMD > $O/ncu_fe.log 2>&1
tail -3 $O/ncu_fe.log
ncu -i $O/r02d_match_best.ncu-rep --page details > $O/r02d_match_best_details.txt 2>/dev/null; grep -n "Duration\|Issue Slots Busy\|highest-utilized\|Registers Per\|Achieved Occ" $O/r02d_match_best_details.txt | head
