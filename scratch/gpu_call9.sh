#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gemm_pm.py tests/test_block.py -m gpu -x -q 2>&1 | tail -30 ) > $OUT/c11_tests_a.log
tail -5 $OUT/c11_tests_a.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/c11_tests.log
tail -3 $OUT/c11_tests.log
for v in 1 0; do
RCOT_PM_TMA=$v timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c11_bench_tma$v.json 2> $OUT/c11_bench_tma$v.err
python - $OUT/c11_bench_tma$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d.get("phases_ms"))
    for k,v in list(d["kernels"].items())[:3]: print("   ",k,v)
except Exception as e:
    print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
timeout 300 python scratch/detail_prof.py 500 > $OUT/c11_detail.txt 2>&1
grep "pm_gemm" $OUT/c11_detail.txt | grep "ks=1" | head -24
