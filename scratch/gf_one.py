"""Fused GDFN middle backward vs the two-kernel form at one level-1 shape (timing; wrap in ncu for a capture)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
B, hid, H, W = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 255, int(sys.argv[2]) if len(sys.argv) > 2 else 128, 0
W = H
u = torch.randn(B, 2 * hid, H, W, device="cuda")
w = torch.randn(2 * hid, 1, 3, 3, device="cuda") / 3
dg = torch.randn(B, hid, H, W, device="cuda")
dw = torch.zeros(2 * hid, 1, 3, 3, device="cuda")
def fused():
    return ops.gdfn_mid_bwd(u, dg, w, dw)
dab = torch.empty_like(u)
def two():
    ops.dwconv(u, w, mode=2, dg=dg, out=dab)
    return ops.dwconv_bwd(u, dab, w, dw)
for fn in (fused, two):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    nb = (2 * u.numel() + dg.numel()) * 4
    print(f"{fn.__name__:6s} hid={hid} {H}x{W}: {t*1000:.1f} us, {nb/t/1e6:.0f} GB/s of the fused form's algorithmic bytes")
