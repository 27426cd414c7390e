#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > $OUT/c19_tests.log
tail -4 $OUT/c19_tests.log
for v in 1 0; do
echo "== RCOT_PK_TMA=$v"
for a in "96 255 0" "96 96 0"; do RCOT_PK_TMA=$v timeout 100 python scratch/pk_one.py $a; done
RCOT_PK_TMA=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c19_bench_$v.json 2> $OUT/c19_bench_$v.err
python - $OUT/c19_bench_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), d["phases_ms"]["T_forward"], d["phases_ms"]["T_backward"], d["kernels"]["pk_gemm"])
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
