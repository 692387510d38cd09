#!/bin/bash
# Selection v3: per-step clocks (RQAE_M3_PROF) and A/B timing
set -u
OUT=gpurun_out; mkdir -p $OUT
RQAE_M3_PROF=1 timeout 120 python tools/bench_select.py --reps 1 2>&1 | tail -3 | tee $OUT/select_prof_r5b.log
timeout 120 python tools/bench_select.py 2>&1 | tail -1 | tee -a $OUT/select_prof_r5b.log
