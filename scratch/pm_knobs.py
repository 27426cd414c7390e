import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
B, C, H, W, N = 32, 48, 128, 128, 144
x = torch.randn(B, C, H, W, device="cuda")
w = torch.randn(N, C, 1, 1, device="cuda") / C ** 0.5
pk = ops.pack_single(w, "fwd")
out = torch.empty(B, N, H, W, device="cuda")
stats = ops.ln_stats(x)
gam, bet = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
def run(debug, ln=True, terms=3, n=5):
    for _ in range(2):
        ops.pm_gemm(x, pk.ptr(0), N, out=out, ln=(stats, gam, bet) if ln else None, debug=debug, terms=terms)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.pm_gemm(x, pk.ptr(0), N, out=out, ln=(stats, gam, bet) if ln else None, debug=debug, terms=terms)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
for dbg, name in [(0, "full"), (7+8+16+32, "bare handshakes"), (7+8+16+32+64, "bare, plain arrives")]:
    print(f"{name:24s} LN {run(dbg):8.1f} us   noLN {run(dbg, ln=False):8.1f} us   terms1 {run(dbg, terms=1):8.1f} us")
