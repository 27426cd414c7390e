#!/usr/bin/env bash
# Round-2 re-check of the whole GPU suite after the bf16-mode test was re-specified (no -x: list every failure).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu --maxfail=8 -rP > gpurun_out/r2j_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2j_tests.log
grep -E "bf16 mode:|passed|failed|^FAILED|tests exit" gpurun_out/r2j_tests.log | tail -12
