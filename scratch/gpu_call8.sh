#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/c8_tests.log
tail -3 $OUT/c8_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c8_bench.json 2> $OUT/c8_bench.err
python - $OUT/c8_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d.get("phases_ms"))
for k,v in list(d["kernels"].items())[:8]: print("   ",k,v)
PY
timeout 300 python scratch/detail_prof.py 500 > $OUT/c8_detail.txt 2>&1
grep -E "ks=[34] mode" $OUT/c8_detail.txt | head -12
