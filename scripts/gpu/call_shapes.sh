#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python scripts/shape_table.py > gpurun_out/r2_shape_table.txt 2>&1
head -70 gpurun_out/r2_shape_table.txt | cut -c1-200
