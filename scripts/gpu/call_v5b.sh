#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for d in 0 3 7; do echo "== debug=$d"; RCOT_GDFN_DEBUG=$d timeout 200 python scripts/prof_gdfn.py 2>&1 | tail -8; done > gpurun_out/r2v5_prof.txt 2>&1
cat gpurun_out/r2v5_prof.txt
