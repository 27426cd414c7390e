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


if __name__ == "__main__":
    main()
