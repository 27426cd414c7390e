import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench, trainer, Net_Restormer as N
from rcot_b200 import ops
B, P = 32, 128
trainer.opt = trainer.parser.parse_args(["--batchSize", str(B), "--patch_size", str(P), "--pairnum", "1000000000", "--no_dump"])
torch.manual_seed(0)
T = N.T_net(decoder=True).cuda(); F = N.F_net(patch_size=P).cuda()
step = trainer._train_step(T, F, "RMSprop")
host = bench.synth_host_batches(1, B, P)
d, t, ids = host[0][1].cuda(), host[0][2].cuda(), host[0][0][1].cuda()
al = torch.rand(B).cuda()
for i in range(2):
    step.iteration(d, t, ids, al, True, 1e-4)
torch.cuda.synchronize()
ops.PROF = ops.Profiler(detail=True)
step.iteration(d, t, ids, al, True, 1e-4)
summ = ops.PROF.summary(); ops.PROF = None
tot = sum(v["ms"] for v in summ.values())
print("total", tot)
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])[:int(sys.argv[1]) if len(sys.argv) > 1 else 45]:
    print(f'{v["ms"]:8.2f} ms {v["ms"]/tot*100:5.1f}% n={v["launches"]:4d} avg={v["ms"]/v["launches"]*1000:8.1f} us  {v["bytes"]/1e9/(v["ms"]/1e3):7.0f} GB/s  {k}')
import re, collections
by = collections.defaultdict(float)
for k, v in summ.items():
    m = re.search(r"\(\d+, \d+, (\d+), \d+\)", k)
    by[m.group(1) if m else "other"] += v["ms"]
print("ms by feature-map height:", {k: round(v, 1) for k, v in sorted(by.items(), key=lambda kv: -kv[1])})
