#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ln_bwd_kernel" -c 3 -o gpurun_out/r2_ln -f python scripts/profile_step.py --warm 0 > gpurun_out/r2nculn.log 2>&1
for i in 0 1 2; do python scripts/ncu_stalls.py gpurun_out/r2_ln.ncu-rep $i; done > gpurun_out/r2_ln_summary.txt 2>&1
cat gpurun_out/r2_ln_summary.txt | cut -c1-120
python scripts/bench_dw.py
