#!/bin/bash
set -u
for c in 1 0; do for t in 512 1184 2368; do echo "RQAE_CLUSTER=$c tokens=$t"; RQAE_CLUSTER=$c timeout 100 python tools/prof_forward.py --dim 3584 --nq 2048 --tokens $t --reps 3 2>&1 | tail -1; done; done
echo "2B small T"; for t in 512 1184; do timeout 100 python tools/prof_forward.py --tokens $t --reps 3 2>&1 | tail -1; done
