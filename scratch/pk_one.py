import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
B, H, W = 32, 128, 128
CA, CB, ln = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
a = torch.randn(B, CA, H, W, device="cuda")
b = torch.randn(B, CB, H, W, device="cuda")
out = torch.zeros(CA, CB, device="cuda")
stats = ops.ln_stats(b)
gam, bet = torch.ones(CB, device="cuda"), torch.zeros(CB, device="cuda")
for _ in range(3):
    ops.pk_gemm(a, b, out, ldo=CB, ln=(stats, gam, bet) if ln else None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.pk_gemm(a, b, out, ldo=CB, ln=(stats, gam, bet) if ln else None)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print(f"pk_gemm CA={CA} CB={CB} ln={ln}: {t*1000:.1f} us  {(a.numel()+b.numel())*4/t/1e6:.0f} GB/s")
