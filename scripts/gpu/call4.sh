#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do echo "debug=$d"; RCOT_GDFN_DEBUG=$d timeout 300 python scripts/bench_gdfn.py 2>&1 | grep "C=96 B=32 128"; done > gpurun_out/r2c4_knobs.txt 2>&1
timeout 1200 python -m pytest tests/test_bench_size.py tests/test_boundary.py tests/test_checkpoint.py tests/test_tester.py tests/test_data.py -m gpu -q > gpurun_out/r2c4_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c4_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c4_launches_b4.csv python bench.py --batch 4 --steps 1 --warmup 3 --graph 0 --no-eager-baseline --no-cpu-baseline --no-profile --no-extra-configs > gpurun_out/r2c4_b4.json 2> gpurun_out/r2c4_b4.err
cat gpurun_out/r2c4_knobs.txt; tail -12 gpurun_out/r2c4_tests.log; wc -l gpurun_out/r2c4_launches_b4.csv
