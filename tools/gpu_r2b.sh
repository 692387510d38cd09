#!/bin/bash
# Round 2, second visit (1 GPU): fused hook tests, host-pipeline probe (debug), quick bench for the new extras.
set -u
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest (hook, forward_host) ==="
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "hook or forward_host" 2>&1 | tail -15 | tee $OUT/pytest_hook_$TAG.log
echo "=== e2e probe N=1 ==="
timeout 400 python tools/e2e_probe.py --tokens 524288 --out $OUT/e2e_probe_n1_$TAG.json > $OUT/e2e_probe_n1_$TAG.log 2>&1; echo "probe exit=$?"
tail -5 $OUT/e2e_probe_n1_$TAG.log
echo "=== bench (short) ==="
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2b.json".replace("r2b","$TAG")))
ex=d["extra"]
print("value",d["value"],"e2e",d["e2e"]["value"],d["e2e"]["int32_codes"]["value"])
print("hook",ex.get("hook_512"))
print("stock",ex.get("stock_pytorch_gpu"))
PY
tail -3 $OUT/bench_$TAG.err
