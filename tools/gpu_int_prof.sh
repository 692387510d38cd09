#!/bin/bash
# ncu capture of the intensity GEMM.  usage: gpurun -- 'bash tools/gpu_int_prof.sh tag [tokens]'
set -u
TAG=${1:-ip}
TOK=${2:-65536}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/bench_intensity.py --tokens $TOK 2>&1 | tail -2 | tee $OUT/bench_intensity_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/int_launches_$TAG.csv \
  python tools/bench_intensity.py --tokens $TOK --reps 1 > $OUT/int_launches_$TAG.log 2>&1
grep -E "rq_intensity|int_" $OUT/int_launches_$TAG.csv | awk -F'","' '{print $5, $NF}' | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_intensity -s 2 -c 1 -f \
  -o $OUT/prof_int_$TAG python tools/bench_intensity.py --tokens $TOK --reps 1 > $OUT/ncu_int_$TAG.log 2>&1
tail -2 $OUT/ncu_int_$TAG.log
