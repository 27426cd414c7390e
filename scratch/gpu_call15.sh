#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $OUT/c15_tests.log
tail -2 $OUT/c15_tests.log
for d in 0 1 2 3 0; do
RCOT_PM_DEBUG=$d timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c15_bench_d$d.json 2> $OUT/c15_bench_d$d.err
python - $OUT/c15_bench_d$d.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), round(d["phases_ms"]["T_forward"],2), round(d["phases_ms"]["T_backward"],2), d["kernels"]["pm_gemm"]["ms"], d["kernels"]["pk_gemm"]["ms"])
PY
done
timeout 300 python scratch/detail_prof.py 500 > $OUT/c15_detail.txt 2>&1
grep -E "ks=3" $OUT/c15_detail.txt | grep pk_gemm | head -8
