#!/bin/bash
set -u
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus $N --steps 2 --warmup 3 > $OUT/bench_n${N}_r2k.json 2> $OUT/bench_n${N}_r2k.err; echo "bench exit=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_n${N}_r2k.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],d["e2e"]["int32_codes"]["value"],d["e2e"]["code_transfer"])
print("c3",d["extra"].get("config3_9b"))
print("c4",d["extra"].get("config4_mining"))
PY
tail -3 $OUT/bench_n${N}_r2k.err
