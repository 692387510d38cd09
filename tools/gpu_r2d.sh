#!/bin/bash
# Round 2: selection kernel v2 (lane-private histograms) -- parity, A/B timing, ncu; bench; then a 2-GPU bench
# (exchange timing of the sharded mining path).
set -u
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest feature + capi ==="
timeout 900 python -m pytest tests/test_feature_gpu.py tests/test_search_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_feat_$TAG.log
echo "=== select A/B ==="
for v in 1 0; do RQAE_MINE_V1=$v timeout 300 python tools/bench_select.py 2>&1 | tail -4; done | tee $OUT/select_ab_$TAG.log
echo "=== ncu select v2 ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_mine2 -s 1 -c 1 -f \
  -o $OUT/prof_mine2_$TAG python tools/bench_select.py --rows 1184 --reps 1 > $OUT/ncu_mine2_$TAG.log 2>&1
tail -2 $OUT/ncu_mine2_$TAG.log
echo "=== e2e probe N=1 (quick) ==="
timeout 400 python tools/e2e_probe.py --tokens 524288 --out $OUT/e2e_probe_n1_$TAG.json > $OUT/e2e_probe_n1_$TAG.log 2>&1; echo "probe exit=$?"
tail -1 $OUT/e2e_probe_n1_$TAG.log | cut -c1-1500
echo "=== bench N=1 ==="
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit=$?"
tail -3 $OUT/bench_$TAG.err
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
ex=d["extra"]
print("value",d["value"],"e2e",d["e2e"]["value"],d["e2e"]["int32_codes"]["value"],d["e2e"]["code_transfer"])
print("hook",ex.get("hook_512"))
print("mining",{k:v for k,v in ex.get("mining",{}).items() if k!="cpu_port"})
print("c4",ex.get("config4_mining"))
PY
