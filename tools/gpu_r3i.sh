#!/bin/bash
# Round-2 second evidence visit: whole -m gpu suite, sanitizer, bench line, launch list, ncu of the intensity GEMM and the search GEMM
set -u
TAG=${1:-r4a}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest -m gpu ==="
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
echo "=== sanitizer ==="
bash tools/sanitize.sh > $OUT/sanitize_$TAG.log 2>&1; grep -E "===|passed|failed|ERROR SUMMARY|error" $OUT/sanitize_$TAG.log | tail -30
echo "=== bench ==="
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit=$?"
tail -3 $OUT/bench_$TAG.err
echo "=== ncu full: intensity GEMM ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_intensity -s 2 -c 1 -f \
  -o $OUT/prof_int_$TAG python tools/bench_intensity.py --tokens 131072 --reps 1 > $OUT/ncu_int_$TAG.log 2>&1
tail -2 $OUT/ncu_int_$TAG.log
