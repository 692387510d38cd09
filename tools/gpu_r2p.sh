#!/bin/bash
# Round-2 evidence visit (1 GPU): whole -m gpu suite, the bench line, the ncu launch list of the same bench command,
# full ncu captures of the forward kernel (2B instantiation) and of the selection kernel.
set -u
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest -m gpu ==="
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "=== bench ==="
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit=$?"
tail -3 $OUT/bench_$TAG.err
echo "=== ncu launch list (same bench command, short) ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --tokens 131072 --no-e2e --no-cpu-baseline --no-parity \
  --tokens-9b 16384 --tokens-mining 131072 > $OUT/bench_under_ncu_$TAG.log 2>&1
tail -2 $OUT/bench_under_ncu_$TAG.log | cut -c1-200
echo "=== ncu full: forward ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_forward -s 1 -c 1 -f \
  -o $OUT/prof_fwd_$TAG python tools/prof_forward.py --tokens 9472 --reps 1 > $OUT/ncu_fwd_$TAG.log 2>&1
tail -2 $OUT/ncu_fwd_$TAG.log
echo "=== ncu full: selection v2 ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_mine2 -s 1 -c 1 -f \
  -o $OUT/prof_mine2_$TAG python tools/bench_select.py --rows 2368 --reps 1 > $OUT/ncu_mine2_$TAG.log 2>&1
tail -2 $OUT/ncu_mine2_$TAG.log
ls -la $OUT | tail -12
