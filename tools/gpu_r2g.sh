#!/bin/bash
set -u
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest feature ==="
timeout 900 python -m pytest tests/test_feature_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "=== select A/B: v1, v2 8 warps x 2 CTAs, v2 16 warps ==="
RQAE_MINE_V1=1 timeout 300 python tools/bench_select.py 2>&1 | tail -1
timeout 300 python tools/bench_select.py 2>&1 | tail -1
RQAE_B200_LIB=$PWD/rqae_b200/librqae_b200_m16.so timeout 300 python tools/bench_select.py 2>&1 | tail -1
timeout 300 python tools/bench_select.py --rows 1792 --n 2097152 --reps 2 2>&1 | tail -1
RQAE_B200_LIB=$PWD/rqae_b200/librqae_b200_m16.so timeout 300 python tools/bench_select.py --rows 1792 --n 2097152 --reps 2 2>&1 | tail -1
