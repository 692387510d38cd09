#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the same bench command and one
# full ncu capture of the forward kernel.  Outputs land in gpurun_out/ (merged back by gpurun).
#   usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1

echo "=== smoke ==="
timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -5 || { echo "SMOKE FAILED/HUNG"; exit 1; }
echo "=== pytest -m gpu ==="
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log

echo "=== bench ==="
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 4000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err

echo "=== decode / forward timing ==="
timeout 300 python tools/prof_forward.py --tokens 262144 --decode --reps 2 2>&1 | tee $OUT/timing_$TAG.log

echo "=== ncu launch list (same bench command, short) ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --tokens 131072 --no-e2e --no-cpu-baseline \
  > $OUT/bench_under_ncu_$TAG.log 2>&1
tail -3 $OUT/bench_under_ncu_$TAG.log | cut -c1-300

echo "=== ncu full capture of the forward kernel ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_forward -s 1 -c 1 -f \
  -o $OUT/prof_fwd_$TAG python tools/prof_forward.py --tokens 9472 --reps 1 > $OUT/ncu_fwd_$TAG.log 2>&1
tail -3 $OUT/ncu_fwd_$TAG.log
echo "=== DRAM traffic of one bench-sized forward launch (1Mi tokens) ==="
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rq_forward -s 1 -c 1 --csv \
  --log-file $OUT/traffic_$TAG.csv python tools/prof_forward.py --tokens 1048576 --reps 1 > $OUT/traffic_$TAG.log 2>&1
python tools/traffic_json.py $OUT/traffic_$TAG.csv 1048576 > $OUT/forward_traffic_$TAG.json; cat $OUT/forward_traffic_$TAG.json
echo "=== ncu full capture of the decode kernel ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_decode -s 1 -c 1 -f \
  -o $OUT/prof_dec_$TAG python tools/prof_forward.py --tokens 9472 --reps 1 --decode > $OUT/ncu_dec_$TAG.log 2>&1
tail -3 $OUT/ncu_dec_$TAG.log
ls -la $OUT
echo "=== ncu full capture of the tensor-core decode (rq_intensity_kernel<1>) ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_intensity -s 2 -c 1 -f \
  -o $OUT/prof_dectc_$TAG python tools/prof_forward.py --tokens 65536 --reps 1 --decode > $OUT/ncu_dectc_$TAG.log 2>&1
tail -3 $OUT/ncu_dectc_$TAG.log
