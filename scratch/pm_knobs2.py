import os, sys, torch
sys.path.insert(0, os.getcwd())
from rcot_b200 import ops
B, H, W = 32, 128, 128
def bench(C, N, ln, res, dbg, terms=3, iters=20):
    x = torch.randn(B, C, H, W, device="cuda")
    w = torch.randn(N, C, 1, 1, device="cuda") / C ** 0.5
    pk = ops.pack_single(w, "fwd")
    out = torch.empty(B, N, H, W, device="cuda")
    r = torch.randn(B, N, H, W, device="cuda") if res else None
    stats = ops.ln_stats(x)
    gam, bet = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    def call():
        ops.pm_gemm(x, pk.ptr(0), N, out=out, ln=(stats, gam, bet) if ln else None, residual=r, debug=dbg, terms=terms)
    call(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                call()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1000
    byts = (x.numel() + out.numel() + (r.numel() if res else 0)) * 4
    return t, byts / t / 1e3
for (C, N, ln, res) in [(48, 144, True, False), (96, 96, False, True), (96, 510, True, False), (255, 96, False, True), (96, 288, True, False)]:
    row = []
    for dbg, name in [(0, "full")]:
        t, gbs = bench(C, N, ln, res, dbg)
        row.append(f"{name}={t:6.1f}us")
    t, gbs = bench(C, N, ln, res, 0)
    t1, _ = bench(C, N, ln, res, 0, terms=1)
    print(f"K={C:4d} N={N:4d} ln={int(ln)} res={int(res)}: " + " ".join(row) + f" | full {gbs:6.0f} GB/s | terms1 {t1:6.1f}us")
