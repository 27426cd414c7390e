#!/usr/bin/env python
"""Micro-benchmark of the 1x1 pixel-as-M GEMMs at the level-1/2 shapes of one training step (x count per step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import engine, ops  # noqa: E402
from scripts.bench_gdfn import timeit  # noqa: E402

B = 32
tot = 0.0
for (Cin, N, H, ln, res, n) in [(96, 96, 128, False, True, 40), (96, 510, 128, True, False, 16), (510, 96, 128, False, False, 16),
                                (255, 96, 128, False, True, 16), (96, 288, 128, True, False, 16), (96, 255, 128, False, False, 16),
                                (192, 192, 128, False, False, 16), (288, 96, 128, False, False, 16), (96, 96, 64, False, True, 56),
                                (96, 510, 64, True, False, 26), (510, 96, 64, False, False, 26), (192, 192, 32, False, True, 54),
                                (192, 1020, 32, True, False, 26), (48, 48, 128, False, True, 16), (48, 254, 128, True, False, 8)]:
    g = torch.Generator().manual_seed(0)
    w = torch.randn(N, Cin, 1, 1, generator=g) / Cin ** 0.5
    sd = {"w.weight": w, "n.body.weight": torch.ones(Cin), "n.body.bias": torch.zeros(Cin)}
    ps = engine.ParamSet(sd, "cuda")
    ps.add_pack("w.weight", "fwd")
    ps.finalize()
    xs = [torch.randn(B, Cin, H, H, device="cuda") for _ in range(2)]
    rs = [torch.randn(B, N, H, H, device="cuda") for _ in range(2)] if res else None
    sts = [ops.ln_stats(x) for x in xs] if ln else None
    i = [0]

    def run():
        i[0] += 1
        k = i[0] % 2
        return ops.pm_gemm(xs[k], ps.pack("w.weight", "fwd"), N, ln=(sts[k], ps.p["n.body.weight"], ps.p["n.body.bias"]) if ln else None,
                           residual=rs[k] if res else None)
    ms = timeit(run)
    tot += ms * n
    nb = (Cin + N + (N if res else 0)) * B * H * H * 4
    print(f"pm_gemm Cin={Cin} N={N} {H}x{H} ln={int(ln)} res={int(res)}: {ms * 1e3:7.1f} us {nb / 1e9 / (ms / 1e3):6.0f} GB/s   x{n}")
print(f"weighted total of these shapes: {tot:.2f} ms per step")
