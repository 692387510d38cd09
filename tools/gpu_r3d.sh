#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_search_gpu.py -q -x -s -k "tensor_core" 2>&1 | tail -25 | tee $OUT/pytest_search_tc_r3d.log
