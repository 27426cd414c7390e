#!/usr/bin/env python
"""Stand-alone launches of the depthwise kernels at the GDFN level-1 shape (C=96: 510 hidden channels, 128x128, B=32)
for ncu captures: gate forward (mode 1), gate backward (mode 2), fused data + weight gradient (dwconv_bwd), plain (mode 0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import ops  # noqa: E402

B, Cn, H, W = 32, 510, 128, 128
u = torch.randn(B, Cn, H, W, device="cuda")
w = torch.randn(Cn, 1, 3, 3, device="cuda") / 3
dw = torch.zeros_like(w)
dg = torch.randn(B, Cn // 2, H, W, device="cuda")
for _ in range(2):
    g = ops.dwconv(u, w, mode=1)
    dab = ops.dwconv(u, w, mode=2, dg=dg)
    du = ops.dwconv_bwd(u, dab, w, dw)
    q = ops.dwconv(u[:, :288], w[:288], mode=0)
    torch.cuda.synchronize()
