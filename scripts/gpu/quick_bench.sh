#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-extra-configs > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2q_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['phases_ms'], d['gpu_launches'])
for k,v in list(d['kernels'].items())[:9]: print(k, v)"
tail -2 gpurun_out/r2q_bench.err
