#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c8_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c8_tests.log
for d in 0 1 2 3 8; do echo "debug=$d"; RCOT_GDFN_DEBUG=$d timeout 300 python scripts/bench_gdfn.py 2>&1 | grep -E "C=96 B=32 128|C=48"; done > gpurun_out/r2c8_knobs.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
echo "bench exit $?" >> gpurun_out/r2c8_bench.err
grep -E "passed|failed|FAILED" gpurun_out/r2c8_tests.log | tail -15; cat gpurun_out/r2c8_knobs.txt; tail -3 gpurun_out/r2c8_bench.err; tail -c 1500 gpurun_out/r2c8_bench.json
