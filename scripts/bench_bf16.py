#!/usr/bin/env python
"""Micro-benchmark of the byte-heavy kernels with fp32 vs bf16 hidden tensors (C=96, 128x128, B=32)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B, C, H, W = 32, 96, 128, 128
    hid = 255
    x = torch.randn(B, C, H, W, device="cuda")
    stats = ops.ln_stats(x)
    gam, bet = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    w_in = (torch.randn(2 * hid, C, 1, 1, device="cuda") / C ** 0.5)
    w_out = (torch.randn(C, hid, 1, 1, device="cuda") / hid ** 0.5)
    wdw = torch.randn(2 * hid, 1, 3, 3, device="cuda") / 3
    pin, pout = ops.pack_single(w_in, "fwd"), ops.pack_single(w_out, "fwd")
    pin_d, pout_d = ops.pack_single(w_in, "dgrad"), ops.pack_single(w_out, "dgrad")
    dy = torch.randn(B, C, H, W, device="cuda")
    dw = torch.zeros_like(wdw)
    dWi, dWo = torch.zeros(2 * hid, C, device="cuda"), torch.zeros(C, hid, device="cuda")
    for dt in (torch.float32, torch.bfloat16):
        u = ops.pm_gemm(x, pin.ptr(0), 2 * hid, ln=(stats, gam, bet), out_dtype=dt)
        g = ops.dwconv(u, wdw, mode=1)
        dg = ops.pm_gemm(dy, pout_d.ptr(0), hid, out_dtype=dt)
        dab = torch.empty_like(u)
        du = torch.empty_like(u)
        res = {
            "x->u (LN)": timeit(lambda: ops.pm_gemm(x, pin.ptr(0), 2 * hid, ln=(stats, gam, bet), out_dtype=dt)),
            "dw gate fwd": timeit(lambda: ops.dwconv(u, wdw, mode=1)),
            "g->y (+res)": timeit(lambda: ops.pm_gemm(g, pout.ptr(0), C, residual=x, stats_out=True)),
            "dy->dg": timeit(lambda: ops.pm_gemm(dy, pout_d.ptr(0), hid, out_dtype=dt)),
            "dw gate bwd": timeit(lambda: ops.dwconv(u, wdw, mode=2, dg=dg, out=dab)),
            "dW_o (pk)": timeit(lambda: ops.pk_gemm(dy, g, dWo, ldo=hid)),
            "dw bwd2": timeit(lambda: ops.dwconv_bwd(u, dab, wdw, dw)),
            "dW_in (pk_mm)": timeit(lambda: ops.pk_gemm(du, x, dWi, ldo=C, ln=(stats, gam, bet))),
            "du->dz": timeit(lambda: ops.pm_gemm(du, pin_d.ptr(0), C)),
        }
        print(str(dt), " ".join(f"{k}: {v:.3f}" for k, v in res.items()), " total %.3f ms" % sum(res.values()))


if __name__ == "__main__":
    main()
