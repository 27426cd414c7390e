#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gdfn_fused.py -m gpu -q -x -s > gpurun_out/r2c3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c3_tests.log
timeout 600 python scripts/bench_gdfn.py > gpurun_out/r2c3_gdfn.txt 2>&1
tail -25 gpurun_out/r2c3_tests.log; cat gpurun_out/r2c3_gdfn.txt
