#!/bin/bash
set -u
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest feature ==="
timeout 900 python -m pytest tests/test_feature_gpu.py tests/test_search_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_feat_$TAG.log
echo "=== select A/B ==="
for v in 1 0; do RQAE_MINE_V1=$v timeout 300 python tools/bench_select.py 2>&1 | tail -1; done | tee $OUT/select_ab_$TAG.log
RQAE_MINE_V1=0 timeout 300 python tools/bench_select.py --rows 1792 --n 2097152 --reps 2 2>&1 | tail -1 | tee -a $OUT/select_ab_$TAG.log
RQAE_MINE_V1=1 timeout 300 python tools/bench_select.py --rows 1792 --n 2097152 --reps 2 2>&1 | tail -1 | tee -a $OUT/select_ab_$TAG.log
echo "=== ncu select v2 ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_mine2 -s 1 -c 1 -f \
  -o $OUT/prof_mine2_$TAG python tools/bench_select.py --rows 1184 --reps 1 > $OUT/ncu_mine2_$TAG.log 2>&1
tail -2 $OUT/ncu_mine2_$TAG.log
