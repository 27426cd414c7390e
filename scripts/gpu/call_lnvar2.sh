#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_elem.py tests/test_lnb.py tests/test_block.py -m gpu -q -x 2>&1 | tail -3
{
for v in 0 1 2; do echo "== RCOT_LN_BWD_VAR=$v"; RCOT_LN_BWD_VAR=$v timeout 120 python scripts/bench_ln.py 2>&1 | tail -5; done
} > gpurun_out/r2_lnvar2.txt 2>&1
cat gpurun_out/r2_lnvar2.txt
