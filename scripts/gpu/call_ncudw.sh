#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dw_" -s 4 -c 4 -o gpurun_out/r2_dw -f python scripts/ncu_dw.py > gpurun_out/r2ncudw.log 2>&1
for i in 0 1 2 3; do python scripts/ncu_stalls.py gpurun_out/r2_dw.ncu-rep $i; done > gpurun_out/r2_dw_summary.txt 2>&1
cat gpurun_out/r2_dw_summary.txt | cut -c1-120
