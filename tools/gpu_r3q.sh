#!/bin/bash
# final evidence: launch list of the bench command, ncu --set full of the search GEMM and of the rows kernel
set -u
TAG=${1:-r3q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --tokens 131072 --no-e2e --no-cpu-baseline --no-parity \
  --tokens-9b 16384 --tokens-mining 131072 > $OUT/bench_under_ncu_$TAG.log 2>&1
tail -1 $OUT/bench_under_ncu_$TAG.log | cut -c1-160
cat > /tmp/tc_one.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from rqae_b200 import RQAE
from rqae_b200.search import IntensityEngine, SERVER_LAYERS
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = RQAE(dim=2304, num_quantizers=1024).eval().to(dev)
N = 4096
codes = torch.randint(0, 625, (N, 127, 1024), generator=torch.Generator(device=dev).manual_seed(77), device=dev, dtype=torch.int16)
eng = IntensityEngine(model, codes, precision="tc")
for _ in range(2):
    for r, l in eng.find_examples(idx=7):
        pass
torch.cuda.synchronize()
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rq_intensity_kernel -s 1 -c 1 -f \
  -o $OUT/prof_searchtc_$TAG python /tmp/tc_one.py > $OUT/ncu_searchtc_$TAG.log 2>&1
tail -2 $OUT/ncu_searchtc_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_rows -s 1 -c 1 -f \
  -o $OUT/prof_searchrows_$TAG python /tmp/tc_one.py > $OUT/ncu_searchrows_$TAG.log 2>&1
tail -2 $OUT/ncu_searchrows_$TAG.log
