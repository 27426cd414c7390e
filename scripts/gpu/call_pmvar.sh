#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{ for v in 4 3 5 6; do echo "== RCOT_PM_RAW=$v"; RCOT_PM_RAW=$v timeout 200 python scripts/bench_pm.py 2>&1 | tail -16; done; } > gpurun_out/r2_pmvar.txt 2>&1
cat gpurun_out/r2_pmvar.txt
