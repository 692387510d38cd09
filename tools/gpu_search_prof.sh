#!/bin/bash
# ncu capture of the search accumulate kernel (last layer range, 511 layers).  usage: gpurun -- 'bash tools/gpu_search_prof.sh tag [sequences]'
set -u
TAG=${1:-sr}
N=${2:-4096}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 ncu --set full --clock-control none --import-source on -k regex:search_accumulate -s 12 -c 1 -f \
  -o $OUT/prof_search_$TAG python tools/search_ncu_driver.py $N > $OUT/ncu_search_$TAG.log 2>&1
tail -2 $OUT/ncu_search_$TAG.log
ls -la $OUT/prof_search_$TAG.ncu-rep
