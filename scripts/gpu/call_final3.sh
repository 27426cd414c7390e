#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2i_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r2i_smoke.log
grep -E "passed|failed" gpurun_out/r2i_tests.log | tail -2; tail -2 gpurun_out/r2i_smoke.log
bash scripts/gpu/call_q.sh
