#!/usr/bin/env python
"""Micro-benchmark: GDFN forward, one fused kernel vs the three-launch path, at the shapes that carry the bytes.
Prints ms per call and GB/s on the ALGORITHMIC bytes of the fused op (read x + write y = 2*B*C*H*W*4 + weights)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rcot_b200 import engine, ops  # noqa: E402


def params(C, g):
    hid = int(C * 2.66)
    r = lambda *s: torch.randn(*s, generator=g)
    return {"b.norm2.body.weight": 1 + 0.2 * r(C), "b.norm2.body.bias": 0.2 * r(C),
            "b.ffn.project_in.weight": r(2 * hid, C, 1, 1) / C ** 0.5, "b.ffn.dwconv.weight": r(2 * hid, 1, 3, 3) / 3,
            "b.ffn.project_out.weight": r(C, hid, 1, 1) / hid ** 0.5}, hid


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    g = torch.Generator().manual_seed(0)
    for (C, B, H, W) in [(48, 32, 128, 128), (96, 32, 128, 128), (96, 32, 64, 64), (96, 8, 256, 256), (96, 4, 128, 128)]:
        sd, hid = params(C, g)
        xs = [torch.randn(B, C, H, W, device="cuda") for _ in range(3)]     # > L2 in rotation at the big shapes
        for x in xs:
            x._rcot_ln_stats = ops.ln_stats(x)
        res = {}
        for fused in (True, False):
            ps = engine.ParamSet(dict(sd), "cuda")
            bs = engine.BlockSpec(ps, "b.", C, 1, has_attn=False)
            if fused:
                ps.add_gdfn("b.", C, hid)
            else:
                ps.gdfn.clear()
            ps.finalize()
            engine.FUSED_GDFN_ALWAYS = fused
            for keep in (False, True):
                i = [0]

                def run():
                    i[0] += 1
                    return engine.gdfn_fwd(bs, xs[i[0] % 3], "b.norm2", True, keep=keep)
                res[(fused, keep)] = timeit(run)
        alg = 2 * B * C * H * W * 4 + 3 * C * hid * 4
        print(f"C={C} B={B} {H}x{W}: fused {res[(True, False)]:.3f} ms ({alg / 1e6 / res[(True, False)]:.0f} GB/s alg), "
              f"fused+save {res[(True, True)]:.3f} ms, unfused {res[(False, False)]:.3f} ms, unfused+keep {res[(False, True)]:.3f} ms")


def main_mdta():
    """MDTA phase 1: one fused kernel vs pm_gemm (x -> pre) + dw_plain + pk_gemm (Gram); algorithmic bytes = read x + write v."""
    g = torch.Generator().manual_seed(1)
    for (C, heads, B, H, W) in [(48, 1, 32, 128, 128), (96, 1, 32, 128, 128), (96, 2, 32, 64, 64), (96, 1, 8, 256, 256), (96, 1, 4, 128, 128)]:
        r = lambda *s: torch.randn(*s, generator=g)
        sd = {"b.norm1.body.weight": 1 + 0.2 * r(C), "b.norm1.body.bias": 0.2 * r(C),
              "b.attn.qkv.weight": r(3 * C, C, 1, 1) / C ** 0.5, "b.attn.qkv_dwconv.weight": r(3 * C, 1, 3, 3) / 3,
              "b.attn.project_out.weight": r(C, C, 1, 1) / C ** 0.5, "b.attn.temperature": torch.ones(heads, 1, 1)}
        xs = [torch.randn(B, C, H, W, device="cuda") for _ in range(3)]
        for x in xs:
            x._rcot_ln_stats = ops.ln_stats(x)
        ps = engine.ParamSet(dict(sd), "cuda")
        engine.BlockSpec(ps, "b.", C, heads, has_ffn=False)
        ps.finalize()
        blob = torch.empty(ops.mdta_p1_blob_bytes(C), dtype=torch.uint8, device="cuda")
        ops.mdta_p1_pack(ps.p["b.attn.qkv.weight"], ps.p["b.attn.qkv_dwconv.weight"], blob)
        c = C // heads
        G = torch.zeros(B, heads, c, c, device="cuda")
        ss = torch.zeros(B, 2 * C, device="cuda")
        res = {}
        i = [0]

        def lnargs(x):
            return (x._rcot_ln_stats, ps.p["b.norm1.body.weight"], ps.p["b.norm1.body.bias"])

        def fused(save):
            def run():
                i[0] += 1
                x = xs[i[0] % 3]
                return ops.mdta_p1(x, blob, heads, G, ss, ln=lnargs(x), save=save)
            return run

        def unfused():
            i[0] += 1
            x = xs[i[0] % 3]
            pre = ops.pm_gemm(x, ps.pack("b.attn.qkv.weight", "fwd"), 3 * C, ln=lnargs(x))
            qkv = ops.dwconv(pre, ps.p["b.attn.qkv_dwconv.weight"], sumsq=ss, nsq=2 * C)
            ops.pk_gemm(qkv[:, :C], qkv[:, C:2 * C], G, ldo=c, per_image=True, groups=heads, out_gs=c * c)
            return qkv
        res["fused"] = timeit(fused(False))
        res["fused+save"] = timeit(fused(True))
        res["unfused"] = timeit(unfused)
        alg = 2 * B * C * H * W * 4 + 3 * C * C * 4
        print(f"MDTA-p1 C={C} h={heads} B={B} {H}x{W}: fused {res['fused']:.3f} ms ({alg / 1e6 / res['fused']:.0f} GB/s alg), "
              f"fused+save {res['fused+save']:.3f} ms, unfused (3 launches) {res['unfused']:.3f} ms")


if __name__ == "__main__":
    if "--mdta" in sys.argv:
        main_mdta()
    else:
        main()
