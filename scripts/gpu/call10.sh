#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python scripts/diag_fnet.py > gpurun_out/r2c10_diag.txt 2>&1
cat gpurun_out/r2c10_diag.txt
