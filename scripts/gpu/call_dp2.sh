#!/usr/bin/env bash
# 2-GPU call: NCCL equivalence test + weak/strong scaling lines at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2dp2_gpus.txt
timeout 900 python -m pytest tests/test_dp_nccl.py -m gpu -q -s > gpurun_out/r2dp2_test.log 2>&1
echo "test exit $?" >> gpurun_out/r2dp2_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2dp2_bench.json 2> gpurun_out/r2dp2_bench.err
echo "bench exit $?" >> gpurun_out/r2dp2_bench.err
tail -6 gpurun_out/r2dp2_test.log; tail -3 gpurun_out/r2dp2_bench.err; tail -c 1200 gpurun_out/r2dp2_bench.json
timeout 900 python -m pytest tests/test_bench_size.py tests/test_bf16_mode.py -m gpu -q > gpurun_out/r2dp2_tests2.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2dp2_tests2.log | tail -5
