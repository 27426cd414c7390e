#!/usr/bin/env python
"""Micro-benchmark of the MDTA small-matrix kernels (attn_fwd, attn_bwd = phase 1 + phase 2) per (C, heads), B=32."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import engine, ops  # noqa: E402
from scripts.bench_gdfn import timeit  # noqa: E402

B = 32
for (C, h) in [(384, 8), (192, 4), (96, 2), (96, 1), (96, 4), (48, 1)]:
    c = C // h
    sc = engine.AttnScratch.get(B, C, h, "cuda")
    sc.G.copy_(torch.randn_like(sc.G) * 30)
    sc.sumsq.copy_(torch.rand_like(sc.sumsq) * 100 + 50)
    sc.P.copy_(torch.randn_like(sc.P))
    temp = torch.ones(h, 1, 1, device="cuda")
    w_out = torch.randn(C, C, device="cuda") / C ** 0.5
    dw, dt = torch.zeros(C, C, device="cuda"), torch.zeros(h, 1, 1, device="cuda")

    def fwd():
        ops.attn_fwd(sc.G, sc.sumsq, temp, w_out, sc.A, sc.Gt, sc.Mpack, sc.MTpack, B, C, h)

    def bwd():
        ops.zero_(sc.dA)
        ops.attn_bwd(sc.P, sc.sumsq, temp, w_out, sc.A, sc.Gt, dw, dt, sc.w12(), B, C, h, sc.dA)
    fwd()
    print(f"attn C={C} heads={h}: fwd {timeit(fwd) * 1e3:6.1f} us   bwd (p1 + p2 + zero) {timeit(bwd) * 1e3:6.1f} us")
