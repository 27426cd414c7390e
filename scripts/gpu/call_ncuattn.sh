#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_p1" -s 29 -c 1 -o gpurun_out/r2_attn1 -f python scripts/profile_step.py --warm 0 > gpurun_out/r2ncuattn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd" -s 16 -c 1 -o gpurun_out/r2_attn2 -f python scripts/profile_step.py --warm 0 >> gpurun_out/r2ncuattn.log 2>&1
{ python scripts/ncu_stalls.py gpurun_out/r2_attn1.ncu-rep 0; python scripts/ncu_summarize.py gpurun_out/r2_attn1.ncu-rep --src 0 --top 22; python scripts/ncu_stalls.py gpurun_out/r2_attn2.ncu-rep 0; python scripts/ncu_summarize.py gpurun_out/r2_attn2.ncu-rep --src 0 --top 16; } > gpurun_out/r2_attn_summary.txt 2>&1
cat gpurun_out/r2_attn_summary.txt | cut -c1-150
