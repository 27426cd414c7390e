#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $OUT/c14_tests.log
tail -2 $OUT/c14_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c14_bench.json 2> $OUT/c14_bench.err
python - $OUT/c14_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d.get("phases_ms"))
for k,v in list(d["kernels"].items())[:4]: print("   ",k,v)
PY
timeout 300 python scratch/detail_prof.py 500 > $OUT/c14_detail.txt 2>&1
grep "pm_gemm" $OUT/c14_detail.txt | grep "ks=1" | head -12
