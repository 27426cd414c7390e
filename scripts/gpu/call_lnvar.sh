#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in 20 21 30 31 40 41; do echo "== RCOT_LN_BWD_VAR=$v"; RCOT_LN_BWD_VAR=$v timeout 120 python scripts/bench_ln.py 2>&1 | tail -5; done > gpurun_out/r2_lnvar.txt 2>&1
cat gpurun_out/r2_lnvar.txt
