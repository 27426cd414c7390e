#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3.py tests/test_tnet.py tests/test_train_step.py tests/test_bench_size.py tests/test_boundary.py -m gpu -q -x 2>&1 | tail -3
bash scripts/gpu/call_q.sh
