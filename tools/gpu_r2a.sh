#!/bin/bash
# Round 2, first GPU visit (1 GPU): smoke, the whole -m gpu suite (new: all 4096 KAT tokens, 9B-width KAT of the
# reference), host-pipeline probe at N=1, the full bench line, ncu of the 9B-width forward instantiation.
#   usage: gpurun --timeout 1500 -- 'bash tools/gpu_r2a.sh [tag]'
set -u
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
lscpu | head -25 > $OUT/lscpu_$TAG.txt; free -g >> $OUT/lscpu_$TAG.txt; nvidia-smi topo -m >> $OUT/lscpu_$TAG.txt 2>&1

echo "=== smoke ==="
timeout 240 python __graft_entry__.py --smoke 2>&1 | tail -5 || { echo "SMOKE FAILED/HUNG"; }
echo "=== pytest -m gpu ==="
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 | tee $OUT/pytest_gpu_$TAG.log

echo "=== e2e probe N=1 ==="
timeout 400 python tools/e2e_probe.py --tokens 524288 --out $OUT/e2e_probe_n1_$TAG.json 2>&1 | tail -3

echo "=== bench ==="
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 6000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err

echo "=== ncu full capture of the 9B-width forward kernel (E=14) ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_forward -s 1 -c 1 -f \
  -o $OUT/prof_fwd9b_$TAG python tools/prof_forward.py --dim 3584 --nq 256 --tokens 7104 --reps 1 > $OUT/ncu_fwd9b_$TAG.log 2>&1
tail -3 $OUT/ncu_fwd9b_$TAG.log
ls -la $OUT
