#!/usr/bin/env bash
# compute-sanitizer memcheck over the unit tests of the kernels added in round 2 (conv3, mdta_p1, make_patches,
# ln_fwd / ln_stats_split); small shapes only, bounded by timeouts.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --print-limit 5"
{ echo "== memcheck tests/test_conv3.py tests/test_data.py (small shapes)"
  timeout 50 $S python -m pytest tests/test_conv3.py tests/test_data.py -m gpu -q -x -k "not 128" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
  echo "== memcheck tests/test_mdta_fused.py (small shapes)"
  timeout 45 $S python -m pytest tests/test_mdta_fused.py -m gpu -q -x -k "not 128" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
} > gpurun_out/sanitizer_r2.txt 2>&1
cat gpurun_out/sanitizer_r2.txt
