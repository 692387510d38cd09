#!/bin/bash
# Short GPU visit: smoke + parity tests + forward/decode timing (no ncu).   usage: gpurun -- 'bash tools/gpu_quick.sh tag [ncu]'
set -u
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "=== smoke ==="
timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -5 || { echo "SMOKE FAILED/HUNG"; exit 1; }
echo "=== pytest -m gpu ==="
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
echo "=== timing ==="
timeout 300 python tools/prof_forward.py --tokens 524288 --decode --reps 2 2>&1 | tee $OUT/timing_$TAG.log
if [ "${2:-}" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_forward -s 1 -c 1 -f \
    -o $OUT/prof_fwd_$TAG python tools/prof_forward.py --tokens 9472 --reps 1 > $OUT/ncu_fwd_$TAG.log 2>&1
  tail -2 $OUT/ncu_fwd_$TAG.log
fi
