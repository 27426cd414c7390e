#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python scripts/bench_bf16.py > gpurun_out/r2c9_bf16.txt 2>&1
cat gpurun_out/r2c9_bf16.txt
