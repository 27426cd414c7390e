#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/dp_nccl_worker.py > gpurun_out/r2dp2e.log 2>&1
echo "exit $?" >> gpurun_out/r2dp2e.log
grep -n "rel-L2\|tail\|weights after\|Assert\|assert\|Error\|DP_NCCL_OK\|exit" gpurun_out/r2dp2e.log | head -20
timeout 600 python -m pytest tests/test_dp_nccl.py -m gpu -q 2>&1 | tail -2
