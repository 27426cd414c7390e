#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 bash scripts/make_profiles.sh r2b > gpurun_out/r2b_profiles.log 2>&1
tail -5 gpurun_out/r2b_profiles.log; head -30 gpurun_out/step_r2b_final_summary.txt
