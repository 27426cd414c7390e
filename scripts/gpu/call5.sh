#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gdfn_fused.py -m gpu -q -x > gpurun_out/r2c5_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c5_tests.log
for d in 0 1 2 3 7; do echo "debug=$d"; RCOT_GDFN_DEBUG=$d timeout 300 python scripts/bench_gdfn.py 2>&1 | grep -E "C=96 B=32 128|C=48"; done > gpurun_out/r2c5_knobs.txt 2>&1
timeout 300 python scripts/bench_gdfn.py > gpurun_out/r2c5_gdfn.txt 2>&1
tail -4 gpurun_out/r2c5_tests.log; cat gpurun_out/r2c5_knobs.txt gpurun_out/r2c5_gdfn.txt
