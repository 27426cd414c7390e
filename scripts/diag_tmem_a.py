#!/usr/bin/env python
"""Bring-up of the tcgen05.mma 'TS' form (A operand resident in TMEM): shape sweep + which partial product matches."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import _lib  # noqa: E402

lib = _lib.lib()
import itertools
for (N, K), VAR in itertools.product([(64, 32), (32, 32), (96, 96), (16, 64), (256, 256), (128, 96), (48, 192)], (0, 1, 2)):
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).bfloat16().float()
    B = torch.randn(N, K, generator=g).bfloat16().float()
    ref = A.double() @ B.double().T
    D = torch.full((128, N), float("nan"), device="cuda")
    Ad, Bd = A.cuda(), B.cuda()          # keep alive: a temporary's block is handed to the next allocation
    rc = lib.rcot_selftest_tmem_a(ctypes.c_void_p(Ad.data_ptr()), ctypes.c_void_p(Bd.data_ptr()),
                                  ctypes.c_void_p(D.data_ptr()), N, K, VAR, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "selftest_tmem_a")
    torch.cuda.synchronize()
    Dc = D.cpu().double()
    err = (Dc - ref).abs().max().item()
    msg = f"N={N} K={K} var={VAR}: max_abs_err {err:.3e} (scale {ref.abs().max().item():.2e})"
    if err > 1e-3:
        # which rows / columns are right?
        rows_ok = ((Dc - ref).abs().max(1).values < 1e-3).nonzero().flatten().tolist()
        cols_ok = ((Dc - ref).abs().max(0).values < 1e-3).nonzero().flatten().tolist()
        msg += f" rows_ok={rows_ok[:8]}..({len(rows_ok)}) cols_ok={cols_ok[:8]}..({len(cols_ok)})"
        for k16 in range(K // 16):
            part = A[:, :16 * (k16 + 1)].double() @ B[:, :16 * (k16 + 1)].double().T
            if (Dc - part).abs().max().item() < 1e-3:
                msg += f" == first {k16 + 1} k16 steps"
    print(msg)
    if err > 1e-3:
        torch.set_printoptions(precision=3, linewidth=200)
        print(" D[:4,:6]\n", Dc[:4, :6], "\n ref[:4,:6]\n", ref[:4, :6])
        # does a wrong row equal some other row of ref / a partial-K product / a product with wrong-B rows?
        for m in (0, 1, 5):
            d = (ref - Dc[m:m + 1]).abs().max(1).values
            print(f"  row {m}: best matching ref row {d.argmin().item()} (err {d.min().item():.2e})")
        # per-row: D[m] = sum_k A'[m,k] B[n,k] -> solve for A' by least squares and compare with A
        Ap = torch.linalg.lstsq(B.double(), Dc[:8].T).solution.T      # [8, K]
        print("  recovered A rows 0..1 (first 8 k):\n", Ap[:2, :8], "\n  true A rows 0..1:\n", A[:2, :8].double())
