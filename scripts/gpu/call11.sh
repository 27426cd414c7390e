#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python scripts/diag_fnet.py 2>&1 | grep -E "^P=|flips" > gpurun_out/r2c11_diag.txt
timeout 600 python scripts/bench_bf16.py > gpurun_out/r2c11_bf16.txt 2>&1
timeout 1800 python -m pytest tests/test_bf16_mode.py tests/test_bench_size.py tests/test_elem.py tests/test_block.py tests/test_gemm_pm.py -m gpu -q > gpurun_out/r2c11_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c11_tests.log
cat gpurun_out/r2c11_diag.txt gpurun_out/r2c11_bf16.txt; grep -E "passed|failed|FAILED" gpurun_out/r2c11_tests.log | tail
