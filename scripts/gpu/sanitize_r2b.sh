#!/usr/bin/env bash
# second sanitizer pass: racecheck on the shared-memory kernels of conv3.cu, memcheck on the fused GDFN (v6) and elem tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{ echo "== racecheck tests/test_conv3.py (small shapes)"
  timeout 30 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_conv3.py -m gpu -q -x -k "not 128" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | head -8
  echo "== memcheck tests/test_gdfn_fused.py tests/test_elem.py (small shapes)"
  timeout 38 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gdfn_fused.py tests/test_elem.py -m gpu -q -x -k "not 128" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
} > gpurun_out/sanitizer_r2b.txt 2>&1
cat gpurun_out/sanitizer_r2b.txt
