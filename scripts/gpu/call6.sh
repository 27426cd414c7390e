#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gdfn_fwd_kernel -s 1 -c 1 -o gpurun_out/r2_gdfn_fused -f python scripts/ncu_gdfn.py > gpurun_out/r2c6_ncu1.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gdfn_fwd_kernel|pm_gemm_kernel|dw_gate_kernel" --csv --log-file gpurun_out/r2c6_traffic.csv python scripts/ncu_gdfn.py > gpurun_out/r2c6_ncu2.log 2>&1
timeout 1500 python -m pytest tests/test_bench_size.py tests/test_checkpoint.py -m gpu -q > gpurun_out/r2c6_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c6_tests.log
tail -5 gpurun_out/r2c6_ncu1.log; cat gpurun_out/r2c6_traffic.csv | tail -30; tail -8 gpurun_out/r2c6_tests.log
