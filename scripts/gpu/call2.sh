#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c2_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench exit $?" >> gpurun_out/r2c2_bench.err
timeout 1500 python scripts/loss_curve.py --out gpurun_out/loss_curve_r2.json > gpurun_out/r2c2_loss.log 2>&1
echo "loss exit $?" >> gpurun_out/r2c2_loss.log
tail -15 gpurun_out/r2c2_tests.log; tail -c 600 gpurun_out/r2c2_bench.json; tail -3 gpurun_out/r2c2_bench.err; tail -3 gpurun_out/r2c2_loss.log
