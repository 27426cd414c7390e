#!/usr/bin/env python
"""Is T_net's training-mode forward reproducible from run to run on ONE GPU (P=32, B=4, tape on)?  Prints the relative L2
distance of each run's output to the first run's; split-K atomics alone give ~1e-7."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import Net_Restormer as N  # noqa: E402
from rcot_b200 import engine  # noqa: E402
from rcot_b200.tnet import TnetProgram  # noqa: E402

P, B = int(os.environ.get("P", "32")), 4
torch.manual_seed(0)
T = N.T_net(decoder=True)
Tp = TnetProgram({k: v.detach().cuda() for k, v in T.named_parameters()}, "cuda")
g = torch.Generator().manual_seed(4)
tgt = torch.rand(B, 3, P, P, generator=g)
deg = (tgt + 0.1 * torch.randn(B, 3, P, P, generator=g)).cuda()
ref = None
errs = []
for it in range(int(os.environ.get("N_IT", "40"))):
    tape = engine.Tape(save_hidden=True)
    out = Tp.forward(deg, tape)
    torch.cuda.synchronize()
    o = out.double().clone()
    del tape
    if ref is None:
        ref = o
        continue
    errs.append(((o - ref).norm() / ref.norm()).item())
errs.sort()
tag = " ".join(f"{k[5:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("RCOT_"))
print(f"[{tag or 'defaults'}] T forward run-to-run rel-L2 over {len(errs)} runs: median {errs[len(errs) // 2]:.2e}, max {errs[-1]:.2e}, "
      f"runs above 1e-5: {sum(e > 1e-5 for e in errs)}")
