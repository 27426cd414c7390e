#!/usr/bin/env bash
# The tests whose bounds come from measured run-to-run spreads (atomics order, sign-like RMSprop steps) + smoke().
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_bench_size.py tests/test_bf16_mode.py::test_full_iteration_bf16_mode -q -m gpu -rP > gpurun_out/r2k_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2k_tests.log
grep -E "^losses|batch-32|bf16 mode:|passed|failed|tests exit" gpurun_out/r2k_tests.log | cut -c1-250
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
