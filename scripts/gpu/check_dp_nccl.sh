#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python -m pytest tests/test_dp_nccl.py -m gpu -q -s 2>&1 | grep -E "grads|weights after|passed|failed|Assertion" | cut -c1-170; echo --; done
