#!/bin/bash
# Intensity GEMM: switch-off experiments (incl. no feature-operand copies) + full ncu capture
set -u
TAG=${1:-r2r}
OUT=gpurun_out
mkdir -p $OUT
DBG_LIST="0 64 65 66 16 80 18 82 90" bash tools/gpu_int_dbg.sh 2>&1 | tee $OUT/int_dbg_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_intensity -s 2 -c 1 -f \
  -o $OUT/prof_int_$TAG python tools/bench_intensity.py --tokens 131072 --reps 1 > $OUT/ncu_int_$TAG.log 2>&1
tail -2 $OUT/ncu_int_$TAG.log
