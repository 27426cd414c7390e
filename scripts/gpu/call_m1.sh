#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mdta_fused.py -m gpu -q -x > gpurun_out/r2m1_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2m1_tests.log
timeout 300 python scripts/bench_gdfn.py --mdta > gpurun_out/r2m1_mdta.txt 2>&1
grep -E "max_err|passed|failed|Error|exit" gpurun_out/r2m1_tests.log | tail -15; cat gpurun_out/r2m1_mdta.txt | tail -8
