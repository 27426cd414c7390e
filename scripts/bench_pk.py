#!/usr/bin/env python
"""Micro-benchmark of the pixel-as-K GEMMs (weight gradients, per-image Grams) at the shapes of one training step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import ops  # noqa: E402
from scripts.bench_gdfn import timeit  # noqa: E402

B = 32
tot = 0.0
# (CA, CB, H, per_image, groups, count per step)
for (CA, CB, H, per_image, groups, n) in [(96, 96, 128, True, 1, 32), (96, 96, 64, True, 2, 56), (192, 192, 32, True, 4, 54),
                                          (48, 48, 128, True, 1, 16), (384, 384, 16, True, 8, 36), (510, 96, 128, False, 1, 16),
                                          (288, 96, 128, False, 1, 16), (96, 255, 128, False, 1, 16), (510, 96, 64, False, 1, 26),
                                          (288, 96, 64, False, 1, 26), (96, 255, 64, False, 1, 26), (1020, 192, 32, False, 1, 26),
                                          (576, 192, 32, False, 1, 26), (192, 510, 32, False, 1, 26), (2042, 384, 16, False, 1, 18),
                                          (254, 48, 128, False, 1, 8), (144, 48, 128, False, 1, 8)]:
    a = [torch.randn(B, CA, H, H, device="cuda") for _ in range(2)]
    b = [torch.randn(B, CB, H, H, device="cuda") for _ in range(2)]
    if per_image:
        c = CA // groups
        out = torch.zeros(B, groups, c, c, device="cuda")
        kw = dict(ldo=c, per_image=True, groups=groups, out_gs=c * c)
    else:
        out = torch.zeros(CA, CB, device="cuda")
        kw = dict(ldo=CB)
    i = [0]

    def run():
        i[0] += 1
        ops.pk_gemm(a[i[0] % 2], b[i[0] % 2], out, **kw)
    ms = timeit(run)
    tot += ms * n
    print(f"pk_gemm A={CA} B={CB} {H}x{H} per_image={int(per_image)} g={groups}: {ms * 1e3:7.1f} us "
          f"{(CA + CB) * B * H * H * 4 / 1e9 / (ms / 1e3):6.0f} GB/s   x{n}")
print(f"weighted total of these shapes: {tot:.2f} ms per step")
