#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c13_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c13_tests.log
timeout 300 python scripts/bench_gdfn.py > gpurun_out/r2c13_gdfn.txt 2>&1
RCOT_LNB_EPILOGUE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-extra-configs > gpurun_out/r2c13_bench_lnb0.json 2> gpurun_out/r2c13_bench_lnb0.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-extra-configs > gpurun_out/r2c13_bench_lnb1.json 2> gpurun_out/r2c13_bench_lnb1.err
grep -E "passed|failed|FAILED" gpurun_out/r2c13_tests.log | tail; cat gpurun_out/r2c13_gdfn.txt; for f in lnb0 lnb1; do python -c "
import json,sys
d=json.loads(open('gpurun_out/r2c13_bench_$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['phases_ms'], {k:(v['launches'],v['ms']) for k,v in d['kernels'].items() if k in ('pm_gemm','ln_bwd','pk_gemm')})
"; done
