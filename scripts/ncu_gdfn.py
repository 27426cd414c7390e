#!/usr/bin/env python
"""One fused / unfused GDFN forward at C=96, 128x128, B=32 (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import engine, ops  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_gdfn import params  # noqa: E402

C, B, H, W = int(os.environ.get("GC", 96)), int(os.environ.get("GB", 32)), 128, 128
g = torch.Generator().manual_seed(0)
sd, hid = params(C, g)
x = torch.randn(B, C, H, W, device="cuda")
x._rcot_ln_stats = ops.ln_stats(x)
for fused in (True, False):
    ps = engine.ParamSet(dict(sd), "cuda")
    bs = engine.BlockSpec(ps, "b.", C, 1, has_attn=False)
    if fused:
        ps.add_gdfn("b.", C, hid)
    else:
        ps.gdfn.clear()
    ps.finalize()
    for _ in range(3):
        y = engine.gdfn_fwd(bs, x, "b.norm2", True, keep=False)
    torch.cuda.synchronize()
