#!/bin/bash
# D-split cluster variant of the forward kernel at 9B width: parity (bit-exact vs the C oracle in the cluster order,
# reference golden under the near-tie protocol), A/B timing against the single-CTA kernel.  Short timeouts: a cluster
# kernel that deadlocks must not hang the box.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "=== parity 9B (cluster) ==="
timeout 180 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "9b or gemma9b" 2>&1 | tail -8
echo "exit=$?"
echo "=== A/B timing 9B encode+recon, 262144 tokens ==="
for c in 1 0; do echo "RQAE_CLUSTER=$c"; RQAE_CLUSTER=$c timeout 200 python tools/prof_forward.py --dim 3584 --nq 2048 --tokens 262144 --reps 1 2>&1 | tail -1; done
echo "=== small token counts (cluster) ==="
for t in 512 2368 4736; do timeout 100 python tools/prof_forward.py --dim 3584 --nq 2048 --tokens $t --reps 2 2>&1 | tail -1; done
