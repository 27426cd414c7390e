#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/c4_tests.log
tail -3 $OUT/c4_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c4_bench.json 2> $OUT/c4_bench.err
python - $OUT/c4_bench.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d.get("phases_ms"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
timeout 300 python scratch/detail_prof.py 500 > $OUT/c4_detail.txt 2>&1
{ for a in "96 1" "192 4" "384 8"; do timeout 120 python scratch/attn_one.py $a; done; } > $OUT/c4_micro.txt 2>&1
cat $OUT/c4_micro.txt
rm -f $OUT/pkmm3.ncu-rep $OUT/gf3.ncu-rep
