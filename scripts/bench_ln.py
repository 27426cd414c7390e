#!/usr/bin/env python
"""Micro-benchmark of the LayerNorm backward kernel at the shapes that carry its time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import ops  # noqa: E402
from scripts.bench_gdfn import timeit  # noqa: E402

for (B, C, H, W) in [(32, 96, 128, 128), (32, 48, 128, 128), (32, 96, 64, 64), (32, 192, 32, 32), (32, 384, 16, 16)]:
    xs = [torch.randn(B, C, H, W, device="cuda") for _ in range(2)]
    dzs = [torch.randn(B, C, H, W, device="cuda") for _ in range(2)]
    dys = [torch.randn(B, C, H, W, device="cuda") for _ in range(2)]
    st = [ops.ln_stats(x) for x in xs]
    gamma = torch.randn(C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    i = [0]

    def run():
        i[0] += 1
        k = i[0] % 2
        return ops.ln_bwd(dzs[k], xs[k], st[k], gamma, dg, db, dy=dys[k])
    ms = timeit(run)
    print(f"ln_bwd {B}x{C}x{H}x{W}: {ms * 1e3:7.1f} us  {4 * B * C * H * W * 4 / 1e9 / (ms / 1e3):6.0f} GB/s")
