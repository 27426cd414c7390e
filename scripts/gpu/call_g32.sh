#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --graph 1 --steps 8 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r2g32_bench.json 2> gpurun_out/r2g32_bench.err
python -c "
import json,torch
d=json.loads([l for l in open('gpurun_out/r2g32_bench.json') if l.startswith('{')][-1])
print('graph=1:', d['ms_per_step'], d['value'], d['e2e'], d['config']['workload'][-30:], d.get('phases_ms'))"
tail -5 gpurun_out/r2g32_bench.err
nvidia-smi --query-gpu=memory.used --format=csv
