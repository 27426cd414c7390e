#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for i in 2 3; do ( RCOT_ATTN_IPC=$i timeout 300 python -m pytest tests/test_block.py -m gpu -x -q 2>&1 | tail -2 ); done
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 )
for n in 1 2 4 8; do echo "== ipc $n"; for a in "96 1" "192 4" "384 8"; do RCOT_ATTN_IPC=$n timeout 100 python scratch/attn_one.py $a | grep bwd; done; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/c21_bench.json 2> $OUT/c21_bench.err
python - $OUT/c21_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), d["phases_ms"]["T_backward"], d["kernels"]["attn_bwd"])
PY
