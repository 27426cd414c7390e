"""Gate forward / gate backward / fused depthwise backward at the GDFN and MDTA shapes (run with RCOT_DW_ROWS=2|4)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
B = 32
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for hid, H in ((255, 128), (127, 128), (255, 64), (510, 32), (1021, 16)):
    u = torch.randn(B, 2 * hid, H, H, device="cuda"); w = torch.randn(2 * hid, 1, 3, 3, device="cuda") / 3
    dg = torch.randn(B, hid, H, H, device="cuda"); dw = torch.zeros_like(w)
    g = torch.empty(B, hid, H, H, device="cuda"); dab = torch.empty_like(u)
    t1 = timeit(lambda: ops.dwconv(u, w, mode=1, out=g))
    t2 = timeit(lambda: ops.dwconv(u, w, mode=2, dg=dg, out=dab))
    t3 = timeit(lambda: ops.dwconv_bwd(u, dab, w, dw))
    nb = u.numel() * 4
    print(f"rows={os.environ.get('RCOT_DW_ROWS','4')} hid={hid} {H}x{H}: gate fwd {t1*1e3:.0f} us ({1.5*nb/t1/1e6:.0f} GB/s)  "
          f"gate bwd {t2*1e3:.0f} us ({2.5*nb/t2/1e6:.0f} GB/s)  dw_bwd2 {t3*1e3:.0f} us ({3*nb/t3/1e6:.0f} GB/s)")
for Cn, H in ((288, 128), (144, 128), (288, 64), (576, 32), (1152, 16)):
    x = torch.randn(B, Cn, H, H, device="cuda"); w = torch.randn(Cn, 1, 3, 3, device="cuda") / 3
    d = torch.randn(B, Cn, H, H, device="cuda"); dw = torch.zeros_like(w)
    t3 = timeit(lambda: ops.dwconv_bwd(x, d, w, dw))
    print(f"rows={os.environ.get('RCOT_DW_ROWS','4')} MDTA Cn={Cn} {H}x{H}: dw_bwd2 {t3*1e3:.0f} us ({3*x.numel()*4/t3/1e6:.0f} GB/s)")
