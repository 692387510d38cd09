#!/bin/bash
# timing-only A/B of library builds on one box:  gpurun -- 'bash tools/ab_quick.sh build/a.so build/b.so ...'
TOK=${TOKENS:-262144}
for rep in 1 2; do
  for so in "$@"; do
    r=$(RQAE_B200_LIB=$PWD/$so timeout 120 python tools/prof_forward.py --tokens $TOK --reps 2 2>&1 | grep "forward ms" | awk '{print $NF}')
    echo "rep$rep $so $r"
  done
done
