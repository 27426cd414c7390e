#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
for c in "1 1" "0 1" "1 0" "0 0"; do set -- $c; RCOT_DIRECT_CONV3=$1 RCOT_PK_SPLIT=$2 timeout 200 python scripts/diag_fgrad.py 2>&1 | tail -1; done
