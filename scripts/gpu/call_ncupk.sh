#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pk_" -s 3 -c 3 -o gpurun_out/r2_pk -f python scripts/ncu_pk.py > gpurun_out/r2ncupk.log 2>&1
{ for i in 0 1 2; do python scripts/ncu_stalls.py gpurun_out/r2_pk.ncu-rep $i; done; python scripts/ncu_summarize.py gpurun_out/r2_pk.ncu-rep --src 0 --top 25 | tail -27; } > gpurun_out/r2_pk_summary.txt 2>&1
cat gpurun_out/r2_pk_summary.txt | cut -c1-150
