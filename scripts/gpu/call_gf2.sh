#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gdfn_fused.py tests/test_tnet.py tests/test_block.py -m gpu -q -x > gpurun_out/r2gf2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2gf2_tests.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gdfn_fwd_kernel|pm_gemm_kernel|dw_gate_kernel" --csv --log-file gpurun_out/r2gf2_traffic.csv python scripts/ncu_gdfn.py > gpurun_out/r2gf2_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gdfn_fwd_kernel -s 1 -c 1 -o gpurun_out/r2_gdfn_fused_v4 -f python scripts/ncu_gdfn.py > gpurun_out/r2gf2_ncu1.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r2gf2_bench.json 2> gpurun_out/r2gf2_bench.err
tail -3 gpurun_out/r2gf2_tests.log; grep gdfn_fwd gpurun_out/r2gf2_traffic.csv | head -3 | cut -c1-60,200-330; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2gf2_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['c5'], d['c2']['ms_per_step'], d['gpu_launches'])"
