#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/diag_tmem_a.py > gpurun_out/r2_tmema.txt 2>&1
cat gpurun_out/r2_tmema.txt
