#!/usr/bin/env python
"""Bring-up of the tcgen05.mma 'TS' form (A operand resident in TMEM): which bf16 packing matches."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import _lib  # noqa: E402

lib = _lib.lib()
for (N, K) in [(32, 32), (64, 96), (96, 64), (256, 128)]:
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).bfloat16().float()
    B = torch.randn(N, K, generator=g).bfloat16().float()
    ref = A.double() @ B.double().T
    for variant in (0, 1):
        D = torch.full((128, N), float("nan"), device="cuda")
        rc = lib.rcot_selftest_tmem_a(ctypes.c_void_p(A.cuda().data_ptr()), ctypes.c_void_p(B.cuda().data_ptr()),
                                      ctypes.c_void_p(D.data_ptr()), N, K, variant,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "selftest_tmem_a")
        torch.cuda.synchronize()
        err = (D.cpu().double() - ref).abs().max().item()
        print(f"N={N} K={K} variant={variant}: max_abs_err {err:.3e} (scale {ref.abs().max().item():.2e})")
