#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > $OUT/c13_tests.log
tail -4 $OUT/c13_tests.log
RCOT_DW_ROWS=4 timeout 200 python scratch/dw_ab.py 2>&1 | grep -E "MDTA|hid" | awk '{print $2,$3,$4,$(NF-3),$(NF-2),$(NF-1),$NF}' > $OUT/c13_dw.txt; cat $OUT/c13_dw.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c13_bench.json 2> $OUT/c13_bench.err
python - $OUT/c13_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d.get("phases_ms"))
for k,v in list(d["kernels"].items())[:6]: print("   ",k,v)
PY
