#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $OUT/c12_tests.log
tail -2 $OUT/c12_tests.log
timeout 200 python scratch/pk_one.py 510 96 1 > $OUT/c12_pk.txt 2>&1; timeout 200 python scratch/pk_one.py 288 96 1 >> $OUT/c12_pk.txt 2>&1; cat $OUT/c12_pk.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/c12_bench.json 2> $OUT/c12_bench.err
python - $OUT/c12_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d.get("phases_ms"), d.get("cpu_baseline"))
PY
timeout 900 bash scripts/make_profiles.sh r1b > $OUT/c12_profiles.log 2>&1
tail -30 $OUT/c12_profiles.log
