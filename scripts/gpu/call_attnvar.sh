#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_block.py tests/test_mdta_fused.py -m gpu -q -x 2>&1 | tail -2
{
for v in 148 296 592 1184; do echo "== RCOT_ATTN_FWD_CTAS=$v"; RCOT_ATTN_FWD_CTAS=$v timeout 120 python scripts/bench_attn.py 2>&1 | tail -6; done
} > gpurun_out/r2_attnvar.txt 2>&1
cat gpurun_out/r2_attnvar.txt
