#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -s -k "9b or gemma9b" 2>&1 | tail -8
{ echo "# Gemma-2-9B width (d=3584, nq=2048), forward (codes + reconstruction), CUDA events, tools/prof_forward.py: single-CTA units (default) vs the opt-in D-split cluster variant (RQAE_CLUSTER=1)";
for c in 0 1; do for t in 512 1184 2368 4736 262144; do echo -n "RQAE_CLUSTER=$c tokens=$t: "; RQAE_CLUSTER=$c timeout 200 python tools/prof_forward.py --dim 3584 --nq 2048 --tokens $t --reps 2 2>&1 | tail -1; done; done; } | tee $OUT/r2n_cluster_ab.txt
