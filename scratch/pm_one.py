import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
B, H, W = 32, 128, 128
C, N, dbg = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(B, C, H, W, device="cuda")
w = torch.randn(N, C, 1, 1, device="cuda") / C ** 0.5
pk = ops.pack_single(w, "fwd")
out = torch.empty(B, N, H, W, device="cuda")
for _ in range(3):
    ops.pm_gemm(x, pk.ptr(0), N, out=out, debug=dbg)
torch.cuda.synchronize()
