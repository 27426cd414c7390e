#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gdfn_fused.py -m gpu -q -x > gpurun_out/r2v6_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2v6_tests.log
timeout 300 python scripts/bench_gdfn.py > gpurun_out/r2v6_gdfn.txt 2>&1
for d in 1 2 3; do echo "debug=$d"; RCOT_GDFN_DEBUG=$d timeout 200 python scripts/bench_gdfn.py 2>&1 | grep -E "C=96 B=32 128"; done > gpurun_out/r2v6_knobs.txt 2>&1
for d in 0 3; do echo "== debug=$d"; RCOT_GDFN_DEBUG=$d timeout 200 python scripts/prof_gdfn.py 2>&1 | tail -9; done > gpurun_out/r2v6_prof.txt 2>&1
tail -5 gpurun_out/r2v6_tests.log; cat gpurun_out/r2v6_gdfn.txt gpurun_out/r2v6_knobs.txt gpurun_out/r2v6_prof.txt
