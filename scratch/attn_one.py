"""attn_fwd / attn_bwd at one (B, C, heads) configuration (timing; wrap in ncu for a capture)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
B, C, h = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 1
c = C // h
dev = "cuda"
G = torch.randn(B, h, c, c, device=dev)
sumsq = torch.rand(B, 2 * C, device=dev) * 100 + 1
temp = torch.rand(h, 1, 1, device=dev) + 0.5
w_out = torch.randn(C, C, device=dev) / C ** 0.5
A = torch.empty(B, h, c, c, device=dev); Gt = torch.empty_like(A)
pb, pb12 = ops.packed_bytes(C, C), ops.packed_bytes(2 * C, 2 * C)
Mp = torch.zeros(B * pb, dtype=torch.uint8, device=dev); MTp = torch.zeros_like(Mp)
W12 = torch.zeros(B * pb12, dtype=torch.uint8, device=dev)
P = torch.randn(B, C, C, device=dev)
dw = torch.zeros(C, C, device=dev); dt = torch.zeros(h, 1, 1, device=dev)
dA = torch.zeros(B, h, c, c, device=dev)
def fwd(): ops.attn_fwd(G, sumsq, temp, w_out, A, Gt, Mp, MTp, B, C, h)
def bwd(): ops.zero_(dA); ops.attn_bwd(P, sumsq, temp, w_out, A, Gt, dw, dt, W12, B, C, h, dA)
for fn in (fwd, bwd):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"attn_{fn.__name__} C={C} heads={h}: {e0.elapsed_time(e1) * 100:.1f} us")
