#!/bin/bash
# compute-sanitizer over small cases of every kernel (memcheck, then racecheck on the shared-memory heavy ones).
#   usage: gpurun --timeout 900 -- 'bash tools/sanitize.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
echo "=== memcheck: mining path (intensity GEMM, transpose, pack, radix select) ==="
timeout 400 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_feature_gpu.py -q -x \
  -k "small or dtypes or api_shapes or 5000 or 4097 or 257 or 100-100" 2>&1 | tail -6 | tee $OUT/sanitize_memcheck_mining.log
echo "=== memcheck: forward / decode small cases ==="
timeout 400 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -x -k "small_forward or decode" 2>&1 | tail -6 | tee $OUT/sanitize_memcheck_fwd.log
echo "=== memcheck: fused hook (rqae_hook_rmsnorm), small-unit instantiation, host pipeline modes ==="
timeout 400 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -x -k "fused_hook or forward_host or invariance" 2>&1 | tail -6 | tee $OUT/sanitize_memcheck_hook.log
echo "=== racecheck: radix select (lane-private counters: v2) ==="
timeout 400 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_feature_gpu.py -q -x -k "5000 or 4097 or 257 or specials or tiny or two_values" 2>&1 | tail -6 | tee $OUT/sanitize_racecheck_mine.log
echo "=== racecheck: sample-bracketed selection (v3: per-warp candidate regions, shared-memory histograms, fallback list) ==="
timeout 400 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_feature_gpu.py -q -x -k "65536 or short_16384 or heavy_median_tie or few_values" 2>&1 | tail -6 | tee $OUT/sanitize_racecheck_mine3.log
echo "=== memcheck: selection v2 / v3, adversarial rows ==="
timeout 400 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_feature_gpu.py -q -x -k "adversarial and not long_2M" 2>&1 | tail -6 | tee $OUT/sanitize_memcheck_mine2.log
echo "=== memcheck + racecheck: example search (table transpose, accumulate, per-position max) ==="
timeout 400 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_search_gpu.py -q -x -k "golden or single_query or argument or int16" 2>&1 | tail -6 | tee $OUT/sanitize_memcheck_search.log
timeout 300 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_search_gpu.py -q -x -k "single_query" 2>&1 | tail -6 | tee $OUT/sanitize_racecheck_search.log
echo "=== memcheck: tensor-core search (store pack, factor gathers, maxima epilogue, exact rows) and tensor-core decode ==="
timeout 400 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_search_gpu.py tests/test_parity_gpu.py -q -x -k "tensor_core and not 127" 2>&1 | tail -6 | tee $OUT/sanitize_memcheck_tc.log
echo "=== initcheck: example search (both modes), selection ==="
# (not the intensity GEMM: initcheck does not see the writes of TMA tensor stores, so every later read of the output is
#  reported as uninitialised although the values are checked against the goldens by the same tests)
timeout 400 $CS --tool initcheck --error-exitcode 9 python -m pytest tests/test_search_gpu.py -q -x -k "tensor_core or single_query or int16" 2>&1 | tail -4 | tee $OUT/sanitize_initcheck_search.log
timeout 400 $CS --tool initcheck --error-exitcode 9 python -m pytest tests/test_feature_gpu.py -q -x -k "4097 or 257 or 5000 or short_16384 or zero_centered or specials" 2>&1 | tail -4 | tee $OUT/sanitize_initcheck_select.log
