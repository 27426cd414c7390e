#!/usr/bin/env python
"""Micro-benchmark of the depthwise backward kernel (data + weight gradient) at the shapes that carry its time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import ops  # noqa: E402
from scripts.bench_gdfn import timeit  # noqa: E402

for (B, Cn, H, W) in [(32, 510, 128, 128), (32, 288, 128, 128), (32, 510, 64, 64), (32, 254, 128, 128), (32, 1020, 32, 32)]:
    us = [torch.randn(B, Cn, H, W, device="cuda") for _ in range(2)]
    ds = [torch.randn(B, Cn, H, W, device="cuda") for _ in range(2)]
    w = torch.randn(Cn, 1, 3, 3, device="cuda") / 3
    dw = torch.zeros_like(w)
    i = [0]

    def run():
        i[0] += 1
        return ops.dwconv_bwd(us[i[0] % 2], ds[i[0] % 2], w, dw)
    ms = timeit(run)
    print(f"dwconv_bwd {B}x{Cn}x{H}x{W}: {ms * 1e3:7.1f} us  {3 * B * Cn * H * W * 4 / 1e9 / (ms / 1e3):6.0f} GB/s")

for (B, Cn, H, W) in [(32, 510, 128, 128), (32, 254, 128, 128), (32, 510, 64, 64), (32, 1020, 32, 32)]:
    us = [torch.randn(B, Cn, H, W, device="cuda") for _ in range(2)]
    dgs = [torch.randn(B, Cn // 2, H, W, device="cuda") for _ in range(2)]
    w = torch.randn(Cn, 1, 3, 3, device="cuda") / 3
    i = [0]

    def fwd():
        i[0] += 1
        return ops.dwconv(us[i[0] % 2], w, mode=1)

    def bwd():
        i[0] += 1
        return ops.dwconv(us[i[0] % 2], w, mode=2, dg=dgs[i[0] % 2])
    m1, m2 = timeit(fwd), timeit(bwd)
    n = B * Cn * H * W * 4
    print(f"gate fwd {B}x{Cn}x{H}x{W}: {m1 * 1e3:7.1f} us {1.5 * n / 1e9 / (m1 / 1e3):6.0f} GB/s   "
          f"gate bwd: {m2 * 1e3:7.1f} us {2.5 * n / 1e9 / (m2 / 1e3):6.0f} GB/s")
