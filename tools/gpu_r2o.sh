#!/bin/bash
set -u
for lib in "" "$PWD/rqae_b200/librqae_b200_noskew.so"; do
  echo "=== lib=${lib:-default (skew 350)} ==="
  for t in 512 1184 2368 65536; do echo -n "2B tokens=$t: "; RQAE_B200_LIB=$lib timeout 100 python tools/prof_forward.py --tokens $t --reps 2 2>&1 | tail -1; done
  for t in 1184 65536; do echo -n "9B cluster tokens=$t: "; RQAE_CLUSTER=1 RQAE_B200_LIB=$lib timeout 100 python tools/prof_forward.py --dim 3584 --nq 2048 --tokens $t --reps 2 2>&1 | tail -1; done
done
