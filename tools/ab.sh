#!/bin/bash
# A/B timing of library builds on ONE box: usage  gpurun -- 'bash tools/ab.sh build/a.so build/b.so ...'
# (build variants with: nvcc $FLAGS -DXXX -o build/x.so rqae_b200/csrc/rqae_capi.cu)
TOK=${TOKENS:-524288}
for so in "$@"; do
  echo "== parity $so: $(RQAE_B200_LIB=$PWD/$so timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k '2b_kat or lockstep or small_forward or boundaries' 2>&1 | tail -1)"
done
for rep in 1 2 3; do
  for so in "$@"; do
    r=$(RQAE_B200_LIB=$PWD/$so timeout 300 python tools/prof_forward.py --tokens $TOK --reps 2 2>&1 | grep "forward ms" | awk '{print $NF}')
    echo "rep$rep $so $r"
  done
done
