#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/c5_tests.log
tail -3 $OUT/c5_tests.log
{ RCOT_DW_ROWS=2 timeout 200 python scratch/dw_ab.py; RCOT_DW_ROWS=4 timeout 200 python scratch/dw_ab.py; timeout 100 python scratch/conv_ab.py; } > $OUT/c5_micro.txt 2>&1
cat $OUT/c5_micro.txt
for r in 2 4; do
RCOT_DW_ROWS=$r timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-profile > $OUT/c5_bench_rows$r.json 2> $OUT/c5_bench_rows$r.err
python - $OUT/c5_bench_rows$r.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d.get("phases_ms"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
