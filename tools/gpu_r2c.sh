#!/bin/bash
# Round 2, 8-GPU visit: where the end-to-end path loses its scaling (probe), and the whole bench line at N=8
# (configs[3] / configs[4] extras on every rank).  Keep it short: charged 8x.
set -u
TAG=${1:-r2c}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" >> $OUT/topo_$TAG.txt; free -g >> $OUT/topo_$TAG.txt
echo "=== pytest (new tests, 1 GPU) ==="
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "hook or forward_host or invariance or max_layers" 2>&1 | tail -4
echo "=== e2e probe N=$N ==="
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  tools/e2e_probe.py --tokens 262144 --out $OUT/e2e_probe_n${N}_$TAG.json > $OUT/e2e_probe_n${N}_$TAG.log 2>&1; echo "probe exit=$?"
tail -4 $OUT/e2e_probe_n${N}_$TAG.log | cut -c1-2500
echo "=== bench N=$N ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus $N --steps 2 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench exit=$?"
tail -c 3000 $OUT/bench_n${N}_$TAG.json; tail -5 $OUT/bench_n${N}_$TAG.err
