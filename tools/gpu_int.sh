#!/bin/bash
# GPU visit for the feature-intensity path: diagnostics, parity tests, timing.  usage: gpurun -- 'bash tools/gpu_int.sh tag'
set -u
TAG=${1:-i}
OUT=gpurun_out
mkdir -p $OUT
echo "=== int_debug ==="
timeout 300 python tools/int_debug.py 2>&1 | tail -60 | tee $OUT/int_debug_$TAG.log
echo "=== pytest feature ==="
timeout 600 python -m pytest tests/test_feature_gpu.py -q 2>&1 | tail -40 | tee $OUT/pytest_feature_$TAG.log
echo "=== bench intensity ==="
timeout 300 python tools/bench_intensity.py 2>&1 | tail -20 | tee $OUT/bench_intensity_$TAG.log
