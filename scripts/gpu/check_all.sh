#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/final_smoke.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench exit $?" >> gpurun_out/final_bench.err
grep -E "passed|failed" gpurun_out/final_tests.log | tail -2; tail -2 gpurun_out/final_smoke.log; tail -2 gpurun_out/final_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1])
for k in ['value','ms_per_step','e2e','gpu_launches','clocks','phases_ms','roofline_blocks','strong','c2','c4','c3_shape_fp32','recompute_mode','c5','bf16_storage','c3','vs_gpu_eager']:
    print(k, ':', json.dumps(d.get(k))[:420])
"
