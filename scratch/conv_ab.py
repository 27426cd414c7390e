"""Stride-2 4x4 transposed convs (F_net data gradients) through pm_gemm: timing at the F_net shapes."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B = 32
for C, Ho in ((64, 64), (128, 32), (256, 16), (512, 8), (512, 4)):
    w = torch.randn(C, C, 4, 4, device="cuda") / (C * 16) ** 0.5
    dy = torch.randn(B, C, Ho, Ho, device="cuda")
    pk = ops.pack_single(w, "dgrad_tap")
    t = timeit(lambda: ops.pm_gemm(dy, pk.ptr(0), C, ks=4, stride=2, pad=1, mode=1, out_hw=(2 * Ho, 2 * Ho), tap_major=True))
    print(f"dgrad 4x4 s2 C={C} out {2*Ho}x{2*Ho}: {t*1e3:.0f} us")
