import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
B, Cn, H, W = 32, 288, 128, 128
x = torch.randn(B, Cn, H, W, device="cuda")
w = torch.randn(Cn, 1, 3, 3, device="cuda")
d = torch.randn(B, Cn, H, W, device="cuda")
dw = torch.zeros(Cn, 1, 3, 3, device="cuda")
sumsq = torch.zeros(B, 192, device="cuda")
for _ in range(3):
    y = ops.dwconv(x, w, sumsq=sumsq, nsq=192)
    din = ops.dwconv_bwd(x, d, w, dw)
torch.cuda.synchronize()
e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e0.record()
for _ in range(10): y = ops.dwconv(x, w, sumsq=sumsq, nsq=192)
e1.record()
for _ in range(10): din = ops.dwconv_bwd(x, d, w, dw)
e2.record(); torch.cuda.synchronize()
nb = x.numel() * 4
print(f"dwconv mode0: {e0.elapsed_time(e1)*100:.1f} us {2*nb/e0.elapsed_time(e1)*10/1e6:.0f} GB/s; dwconv_bwd: {e1.elapsed_time(e2)*100:.1f} us {3*nb/e1.elapsed_time(e2)*10/1e6:.0f} GB/s")
