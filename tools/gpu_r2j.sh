#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "=== smoke ==="
timeout 240 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== full pytest -m gpu ==="
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_r2j.log
bash tools/sanitize.sh
