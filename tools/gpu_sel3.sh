#!/bin/bash
# Selection v3 (sample-bracketed single pass): tests, per-step clocks, A/B timing against the three-pass kernel
set -u
TAG=${1:-r5}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_feature_gpu.py -x -q -k "select or mining or helper" 2>&1 | tail -8
{ RQAE_M3_PROF=1 timeout 120 python tools/bench_select.py --reps 1 2>&1 | tail -3
  for v in 0 1; do RQAE_MINE_V2=$v timeout 120 python tools/bench_select.py 2>&1 | tail -1; done; } | tee $OUT/select_ab_$TAG.log
