#!/usr/bin/env python
"""Is the F-sub gradient (critic step) reproducible from run to run on ONE GPU?  Repeats the same critic step and prints
the relative L2 distance of each run's flat gradient to the first run's, per switch setting (bisects intermittent errors)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import Net_Restormer as N  # noqa: E402
from rcot_b200.fnet import FnetProgram  # noqa: E402

P, B = 32, 4
torch.manual_seed(0)
F = N.F_net(patch_size=P)
Fp = FnetProgram({k: v.detach().cuda() for k, v in F.named_parameters()}, "cuda", P)
g = torch.Generator().manual_seed(4)
real = torch.rand(B, 3, P, P, generator=g).cuda()
fake = (torch.rand(B, 3, P, P, generator=g)).cuda()
ref = None
worst = 0.0
errs = []
for it in range(int(os.environ.get("N_IT", "60"))):
    Fp.ps.zero_grad()
    Fp.critic_step(real, fake, B)
    torch.cuda.synchronize()
    gcur = Fp.ps.grad.double().clone()
    if ref is None:
        ref = gcur
        continue
    e = ((gcur - ref).norm() / ref.norm()).item()
    errs.append(e)
errs.sort()
print(f"DIRECT_CONV3={os.environ.get('RCOT_DIRECT_CONV3', '1')} PK_SPLIT={os.environ.get('RCOT_PK_SPLIT', '1')}: "
      f"run-to-run rel-L2 of the F-sub gradient over {len(errs)} runs: median {errs[len(errs) // 2]:.2e}, max {errs[-1]:.2e}, "
      f"runs above 1e-4: {sum(e > 1e-4 for e in errs)}")
