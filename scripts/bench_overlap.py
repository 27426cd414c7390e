#!/usr/bin/env python
"""Do the weight-gradient GEMMs overlap with the data-gradient chain when launched on a second stream?
GDFN backward at C=96, 128x128, B=32: (a) everything on one stream, (b) the two pk_gemm launches on a side stream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import engine, ops  # noqa: E402
from scripts.bench_gdfn import params  # noqa: E402

C, B, H, W = 96, 32, 128, 128
g = torch.Generator().manual_seed(0)
sd, hid = params(C, g)
ps = engine.ParamSet(dict(sd), "cuda")
bs = engine.BlockSpec(ps, "b.", C, 1, has_attn=False)
ps.gdfn.clear()
ps.finalize()
x = torch.randn(B, C, H, W, device="cuda")
dy = torch.randn(B, C, H, W, device="cuda")
y, kept = engine.gdfn_fwd(bs, x, "b.norm2", True, keep=True)
stats, u, gg = kept
f = "b.ffn."
side = torch.cuda.Stream()
ln = engine._ln_args(ps, "b.norm2", stats)


def chain(use_side):
    main = torch.cuda.current_stream()
    dg = ops.pm_gemm(dy, ps.pack(f + "project_out.weight", "dgrad"), hid)
    if use_side:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.pk_gemm(dy, gg, ps.g[f + "project_out.weight"].view(C, hid), ldo=hid)
    else:
        ops.pk_gemm(dy, gg, ps.g[f + "project_out.weight"].view(C, hid), ldo=hid)
    dab = ops.dwconv(u, ps.p[f + "dwconv.weight"], mode=2, dg=dg)
    du = ops.dwconv_bwd(u, dab, ps.p[f + "dwconv.weight"], ps.g[f + "dwconv.weight"])
    if use_side:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.pk_gemm(du, x, ps.g[f + "project_in.weight"].view(2 * hid, C), ldo=C, ln=ln)
    else:
        ops.pk_gemm(du, x, ps.g[f + "project_in.weight"].view(2 * hid, C), ldo=C, ln=ln)
    dz = ops.pm_gemm(du, ps.pack(f + "project_in.weight", "dgrad"), C)
    dx = ops.ln_bwd(dz, x, stats, ps.p["b.norm2.body.weight"], ps.g["b.norm2.body.weight"], ps.g["b.norm2.body.bias"], dy=dy, dx=dz)
    if use_side:
        main.wait_stream(side)
    return dx


for use_side in (False, True, False, True):
    for _ in range(3):
        chain(use_side)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        chain(use_side)
    b.record()
    torch.cuda.synchronize()
    print(f"GDFN backward C={C} {H}x{W} B={B}: side stream for the weight gradients = {use_side}: {a.elapsed_time(b) / 10:.3f} ms")
