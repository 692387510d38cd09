#!/bin/bash
# ncu --set full of the sample-bracketed selection kernel (2368 rows = 8 per CTA)
set -u
TAG=${1:-r5e}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_mine3 -s 1 -c 1 -f \
  -o $OUT/prof_mine3_$TAG python tools/bench_select.py --rows 2368 --reps 1 > $OUT/ncu_mine3_$TAG.log 2>&1
tail -2 $OUT/ncu_mine3_$TAG.log
ls -la $OUT/prof_mine3_$TAG.ncu-rep
