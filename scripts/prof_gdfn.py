#!/usr/bin/env python
"""Where do the warps of the fused GDFN forward wait?  Runs one launch of the PROF instantiation (RCOT_GDFN_DEBUG | 16)
at C=96, B=32, 128x128 and prints the cycle counters of CTA 0, per warp role (csrc/gdfn_fused.cu, `tw` sites)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RCOT_GDFN_DEBUG"] = str(int(os.environ.get("RCOT_GDFN_DEBUG", "0")) | 16)
import torch  # noqa: E402

from rcot_b200 import engine, ops  # noqa: E402
from scripts.bench_gdfn import params  # noqa: E402

SITES = {
    "issuer1": ["zbar", "winbar", "dbar", "-", "-", "-", "TOTAL", "-"],
    "issuer2": ["-", "-", "-", "sbar", "ybar", "gbar(wo)", "TOTAL", "-"],
    "drain": ["ubar", "uempty", "-", "-", "-", "-", "TOTAL", "-"],
    "stencil": ["ufull", "wobar", "gbar(g slot)", "ufull(nl) peek", "peek + produce_z", "gbar(epi)", "TOTAL", "epilogue incl. gbar wait"],
}


def main():
    C, B, H, W = int(os.environ.get("GF_C", "96")), 32, 128, 128
    g = torch.Generator().manual_seed(0)
    sd, hid = params(C, g)
    x = torch.randn(B, C, H, W, device="cuda")
    x._rcot_ln_stats = ops.ln_stats(x)
    ps = engine.ParamSet(dict(sd), "cuda")
    bs = engine.BlockSpec(ps, "b.", C, 1, has_attn=False)
    ps.add_gdfn("b.", C, hid)
    ps.finalize()
    engine.FUSED_GDFN_ALWAYS = True
    for _ in range(3):
        engine.gdfn_fwd(bs, x, "b.norm2", True, keep=False)
    torch.cuda.synchronize()
    c = ops.gdfn_profile_read()
    tiles = (B * H * W // 128 + 147) // 148
    print(f"C={C}: CTA 0, {tiles} tiles, {(hid + 15) // 16} slices per tile; cycles (share of the warp's total)")
    for name, w in (("issuer1", 20), ("issuer2", 21), ("drain q0", 16), ("drain q3", 19), ("stencil g0 w0", 0), ("stencil g1 w8", 8), ("stencil g1 w15", 15)):
        tot = float(c[w, 6])
        sites = SITES[name.split()[0]]
        parts = ", ".join(f"{s} {int(c[w, i])} ({100 * c[w, i] / tot:.0f}%)" for i, s in enumerate(sites) if s not in ("-", "TOTAL"))
        print(f"  {name:15s} total {int(tot)}  per tile {tot / tiles:.0f}: {parts}")


if __name__ == "__main__":
    main()
