#!/usr/bin/env bash
# 2-GPU: CUDA-graph replay WITH the NCCL all-reduces captured (what N=4/8 of the scaling run use), DP test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_nccl.py -m gpu -q -s > gpurun_out/r2dp2b_test.log 2>&1
echo "test exit $?" >> gpurun_out/r2dp2b_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --batch 4 --steps 5 --warmup 3 --graph 1 --no-profile > gpurun_out/r2dp2b_graph.json 2> gpurun_out/r2dp2b_graph.err
echo "graph bench exit $?" >> gpurun_out/r2dp2b_graph.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --batch 4 --steps 5 --warmup 3 --graph 0 --no-profile > gpurun_out/r2dp2b_eager.json 2> gpurun_out/r2dp2b_eager.err
tail -4 gpurun_out/r2dp2b_test.log; tail -3 gpurun_out/r2dp2b_graph.err; cut -c1-400 gpurun_out/r2dp2b_graph.json; cut -c1-400 gpurun_out/r2dp2b_eager.json
