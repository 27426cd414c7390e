"""Per-shape CUDA-event table of one adversarial iteration (B=32, P=128): every engine op keyed by its tensor shapes,
sorted by total time.  `python scripts/shape_table.py [--batch 32]` -> stdout."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import Net_Restormer as N
import trainer
from rcot_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--patch", type=int, default=128)
a = ap.parse_args()
trainer.opt = trainer.parser.parse_args(["--batchSize", str(a.batch), "--patch_size", str(a.patch), "--pairnum", "1000000000", "--no_dump"])
torch.manual_seed(0)
T = N.T_net(decoder=True).cuda()
F = N.F_net(patch_size=a.patch).cuda()
step = trainer._train_step(T, F, "RMSprop")
host = bench.synth_host_batches(1, a.batch, a.patch)
d, t, ids = host[0][1].cuda(), host[0][2].cuda(), host[0][0][1].cuda()
al = torch.rand(a.batch).cuda()
for i in range(2):
    step.iteration(d, t, ids, al, True, 1e-4)
torch.cuda.synchronize()
ops.PROF = ops.Profiler(detail=True)
step.iteration(d, t, ids, al, True, 1e-4)
summ = ops.PROF.summary()
ops.PROF = None
tot = sum(v["ms"] for v in summ.values())
print(f"total {tot:.1f} ms over {sum(v['launches'] for v in summ.values())} instrumented calls")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = v["bytes"] / 1e9 / (v["ms"] / 1e3) if v["ms"] > 0 else 0
    print(f"{v['ms']:8.3f} ms {v['launches']:4d} x {v['ms'] / v['launches'] * 1e3:8.1f} us {gbs:7.0f} GB/s  {k}")
