#!/bin/bash
# timing experiments on the intensity GEMM (results are wrong by construction when RQAE_INT_DBG has bits 1-16 set)
# bit 4 (no pause) can dead-lock the epilogue's phase tracking: not in the list
for d in ${DBG_LIST:-0 1 2 8 3 9 10 11}; do
  echo -n "dbg=$d  "
  RQAE_INT_DBG=$d timeout 60 python tools/bench_intensity.py --tokens ${TOKENS:-131072} 2>&1 | tail -1 | cut -c1-120
done
