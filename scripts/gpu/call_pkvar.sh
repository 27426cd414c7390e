#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_block.py tests/test_elem.py tests/test_gemm_pm.py -m gpu -q -x 2>&1 | tail -2
{ for v in 0 1; do echo "== RCOT_PK_SPLIT=$v"; RCOT_PK_SPLIT=$v timeout 200 python scripts/bench_pk.py 2>&1 | tail -19; done; } > gpurun_out/r2_pkvar.txt 2>&1
cat gpurun_out/r2_pkvar.txt
