#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/c20_smoke.log 2>&1; tail -2 $OUT/c20_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/c20_bench.json 2> $OUT/c20_bench.err
python - $OUT/c20_bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"], d.get("phases_ms"), d.get("cpu_baseline"), d["roofline"]["frac"], d["roofline_blocks"]["frac"])
PY
timeout 900 bash scripts/make_profiles.sh r1b > $OUT/c20_profiles.log 2>&1
tail -12 $OUT/c20_profiles.log
