#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 200 python scripts/diag_tfwd.py 2>&1 | tail -1
RCOT_DIRECT_CONV3=0 timeout 200 python scripts/diag_tfwd.py 2>&1 | tail -1
RCOT_ATTN_FWD_CTAS=148 timeout 200 python scripts/diag_tfwd.py 2>&1 | tail -1
RCOT_DW_GATE1_VAR=10 timeout 200 python scripts/diag_tfwd.py 2>&1 | tail -1
RCOT_PK_SPLIT=0 timeout 200 python scripts/diag_tfwd.py 2>&1 | tail -1
RCOT_DIRECT_CONV3=0 RCOT_PK_SPLIT=0 RCOT_DW_GATE1_VAR=10 RCOT_ATTN_FWD_CTAS=148 timeout 200 python scripts/diag_tfwd.py 2>&1 | tail -1
