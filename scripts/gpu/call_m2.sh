#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mdta_fused.py tests/test_gdfn_fused.py tests/test_block.py tests/test_tnet.py tests/test_tester.py tests/test_boundary.py -m gpu -q -x > gpurun_out/r2m2_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2m2_tests.log
grep -E "passed|failed|Error|exit" gpurun_out/r2m2_tests.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r2m2_bench.json 2> gpurun_out/r2m2_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2m2_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['c5'], d['c2']['ms_per_step'], d['gpu_launches'])"
tail -3 gpurun_out/r2m2_bench.err
