#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_nccl.py -m gpu -q -s 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r2dp2d_bench.json 2> gpurun_out/r2dp2d_bench.err
echo "bench exit $?" >> gpurun_out/r2dp2d_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2dp2d_bench.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d.get('phases_ms'), d.get('strong'))"
tail -3 gpurun_out/r2dp2d_bench.err
