#!/usr/bin/env python
"""Stand-alone launches of the pixel-as-K GEMMs at level-1 shapes for ncu: per-image Gram (q k^T), dW of a 1x1 conv."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import ops  # noqa: E402

B, C, H, W = 32, 96, 128, 128
q = torch.randn(B, C, H, W, device="cuda")
k = torch.randn(B, C, H, W, device="cuda")
G = torch.zeros(B, 1, C, C, device="cuda")
q2 = torch.randn(B, C, 64, 64, device="cuda")
k2 = torch.randn(B, C, 64, 64, device="cuda")
G2 = torch.zeros(B, 2, 48, 48, device="cuda")
du = torch.randn(B, 288, H, W, device="cuda")
dW = torch.zeros(288, C, device="cuda")
for _ in range(2):
    ops.pk_gemm(q, k, G, ldo=C, per_image=True, groups=1, out_gs=C * C)
    ops.pk_gemm(q2, k2, G2, ldo=48, per_image=True, groups=2, out_gs=48 * 48)
    ops.pk_gemm(du, q, dW, ldo=C)
    torch.cuda.synchronize()
