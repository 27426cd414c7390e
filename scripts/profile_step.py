"""Runs `warm` + 1 adversarial iterations at the bench configuration; meant to be wrapped in ncu
(see profiles/README.md for the exact command lines)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import Net_Restormer as N
import trainer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--patch", type=int, default=128)
ap.add_argument("--warm", type=int, default=1)
a = ap.parse_args()
trainer.opt = trainer.parser.parse_args(["--batchSize", str(a.batch), "--patch_size", str(a.patch), "--pairnum", "1000000000", "--no_dump"])
torch.manual_seed(0)
T = N.T_net(decoder=True).cuda()
F = N.F_net(patch_size=a.patch).cuda()
step = trainer._train_step(T, F, "RMSprop")
host = bench.synth_host_batches(1, a.batch, a.patch)
d, t, ids = host[0][1].cuda(), host[0][2].cuda(), host[0][0][1].cuda()
al = torch.rand(a.batch).cuda()
from rcot_b200 import ops
for i in range(a.warm + 1):
    l0 = ops.LAUNCHES
    step.iteration(d, t, ids, al, True, 1e-4)
    torch.cuda.synchronize()
    print("launches this step:", ops.LAUNCHES - l0, flush=True)
