#!/usr/bin/env bash
# round-2 call 1: new parity tests + baselines + loss curve
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c1_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c1_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench exit $?" >> gpurun_out/r2c1_bench.err
timeout 1200 python scripts/loss_curve.py --out gpurun_out/loss_curve_r2.json > gpurun_out/r2c1_loss.log 2>&1
echo "loss exit $?" >> gpurun_out/r2c1_loss.log
tail -5 gpurun_out/r2c1_tests.log; tail -c 1500 gpurun_out/r2c1_bench.json; tail -3 gpurun_out/r2c1_loss.log
