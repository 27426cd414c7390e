#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2f_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r2f_smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench exit $?" >> gpurun_out/r2f_bench.err
timeout 1500 bash scripts/make_profiles.sh r2 > gpurun_out/r2f_profiles.log 2>&1
grep -E "passed|failed" gpurun_out/r2f_tests.log | tail -2; tail -2 gpurun_out/r2f_smoke.log; tail -2 gpurun_out/r2f_bench.err; cut -c1-300 gpurun_out/r2f_bench_ref.json; tail -5 gpurun_out/r2f_profiles.log
