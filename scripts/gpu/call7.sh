#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gdfn_fused.py tests/test_bench_size.py tests/test_checkpoint.py -m gpu -q > gpurun_out/r2c7_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c7_tests.log
for d in 0 1 2 3; do echo "debug=$d"; RCOT_GDFN_DEBUG=$d timeout 300 python scripts/bench_gdfn.py 2>&1 | grep -E "C=96 B=32 128|C=48"; done > gpurun_out/r2c7_knobs.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdfn_fwd_kernel -s 1 -c 1 -o gpurun_out/r2_gdfn_fused_v3 -f python scripts/ncu_gdfn.py > gpurun_out/r2c7_ncu1.log 2>&1
tail -8 gpurun_out/r2c7_tests.log; cat gpurun_out/r2c7_knobs.txt
