#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python scripts/bench_overlap.py 2>&1 | tail -6
