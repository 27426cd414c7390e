#!/usr/bin/env python
"""Launches for ncu captures of the fused forward kernels at C=GC (96), 128x128, B=32:
GDFN fused (no save), GDFN fused + save, MDTA phase 1 fused (no save), MDTA phase 1 fused + save, then the unfused
three-launch forms of both (pm_gemm, dw_gate / dw_plain, pm_gemm / pk_gemm)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import engine, ops  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_gdfn import params  # noqa: E402

C, B, H, W = int(os.environ.get("GC", 96)), int(os.environ.get("GB", 32)), 128, 128
heads = 1
g = torch.Generator().manual_seed(0)
sd, hid = params(C, g)
r = lambda *s: torch.randn(*s, generator=g)
sd.update({"b.norm1.body.weight": 1 + 0.2 * r(C), "b.norm1.body.bias": 0.2 * r(C),
           "b.attn.qkv.weight": r(3 * C, C, 1, 1) / C ** 0.5, "b.attn.qkv_dwconv.weight": r(3 * C, 1, 3, 3) / 3,
           "b.attn.project_out.weight": r(C, C, 1, 1) / C ** 0.5, "b.attn.temperature": torch.ones(heads, 1, 1)})
x = torch.randn(B, C, H, W, device="cuda")
x._rcot_ln_stats = ops.ln_stats(x)
ps = engine.ParamSet(dict(sd), "cuda")
bs = engine.BlockSpec(ps, "b.", C, heads)
ps.add_gdfn("b.", C, hid)
ps.add_mdta("b.", C)
ps.finalize()
ln2 = (x._rcot_ln_stats, ps.p["b.norm2.body.weight"], ps.p["b.norm2.body.bias"])
ln1 = (x._rcot_ln_stats, ps.p["b.norm1.body.weight"], ps.p["b.norm1.body.bias"])
G = torch.zeros(B, heads, C // heads, C // heads, device="cuda")
ss = torch.zeros(B, 2 * C, device="cuda")
for rep in range(2):
    for save in (False, True):
        out = ops.gdfn_fwd(x, ps.gdfn_blob("b."), hid, ln=ln2, residual=True, stats_out=True, save=save)
        del out
    for save in (False, True):
        out = ops.mdta_p1(x, ps.mdta_blob("b."), heads, G, ss, ln=ln1, save=save)
        del out
    torch.cuda.synchronize()
# unfused forms
u = ops.pm_gemm(x, ps.pack("b.ffn.project_in.weight", "fwd"), 2 * hid, ln=ln2)
gg = ops.dwconv(u, ps.p["b.ffn.dwconv.weight"], mode=1)
y = ops.pm_gemm(gg, ps.pack("b.ffn.project_out.weight", "fwd"), C, residual=x, stats_out=True)
pre = ops.pm_gemm(x, ps.pack("b.attn.qkv.weight", "fwd"), 3 * C, ln=ln1)
qkv = ops.dwconv(pre, ps.p["b.attn.qkv_dwconv.weight"], sumsq=ss, nsq=2 * C)
ops.pk_gemm(qkv[:, :C], qkv[:, C:2 * C], G, ldo=C // heads, per_image=True, groups=heads, out_gs=(C // heads) ** 2)
torch.cuda.synchronize()
