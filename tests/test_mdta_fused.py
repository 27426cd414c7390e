"""One-kernel MDTA phase 1 (csrc/mdta_fused.cu; Net_Restormer.py:29-41 + norm1) against the fp64 oracle: v, the per-head
Grams q k^T, the row sums of squares of q and k, and the optional saved pre / q,k, for every (C, heads) class the levels
with C in {48, 96} use, ragged tile counts (CTAs that span image boundaries) and the benchmarked 128x128 / 64x64 shapes;
then the whole MDTA forward (phase 1 + small-matrix step + y = x + M v) against the reference module's arithmetic."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(C, heads, g):
    r = lambda *s: torch.randn(*s, generator=g)
    return {"b.norm1.body.weight": 1 + 0.2 * r(C), "b.norm1.body.bias": 0.2 * r(C),
            "b.attn.qkv.weight": r(3 * C, C, 1, 1) / C ** 0.5, "b.attn.qkv_dwconv.weight": r(3 * C, 1, 3, 3) / 3,
            "b.attn.project_out.weight": r(C, C, 1, 1) / C ** 0.5, "b.attn.temperature": 1 + 0.3 * r(heads, 1, 1)}


def _close(name, got, ref, rtol=1e-3, atol=1e-4):
    got = got.detach().cpu().double()
    err = (got - ref).abs()
    tol = atol * max(1.0, ref.abs().max().item()) + rtol * ref.abs()
    bad = (err > tol).sum().item()
    print(f"{name:10s} max_err={err.max().item():.3e} scale={ref.abs().max().item():.3e} bad={bad}")
    assert bad == 0, name


@pytest.mark.parametrize("C,heads,B,H,W", [(48, 1, 2, 8, 16), (96, 2, 1, 16, 32), (96, 4, 3, 24, 16), (48, 1, 2, 128, 128),
                                           (96, 1, 2, 128, 128), (96, 2, 2, 64, 64), (96, 4, 5, 40, 48), (96, 1, 150, 8, 16)])
@pytest.mark.parametrize("ln", [True, False])
def test_mdta_phase1(cuda_lib, C, heads, B, H, W, ln):
    from oracle import restormer_ref as R
    from rcot_b200 import ops
    F = torch.nn.functional
    g = torch.Generator().manual_seed(C + H + W + heads)
    sd = _params(C, heads, g)
    x = torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3
    x64 = x.double()
    sd64 = {k: v.double() for k, v in sd.items()}
    z64 = R.layer_norm_c(x64, sd64["b.norm1.body.weight"], sd64["b.norm1.body.bias"]) if ln else x64
    pre64 = F.conv2d(z64, sd64["b.attn.qkv.weight"])
    qkv64 = F.conv2d(pre64, sd64["b.attn.qkv_dwconv.weight"], padding=1, groups=3 * C)
    c = C // heads
    q64 = qkv64[:, :C].reshape(B, heads, c, H * W)
    k64 = qkv64[:, C:2 * C].reshape(B, heads, c, H * W)
    G64 = q64 @ k64.transpose(-1, -2)
    ss64 = (qkv64[:, :2 * C] ** 2).sum(dim=(2, 3))
    assert ops.mdta_p1_supported(C, H, W, heads)
    xd = x.cuda()
    blob = torch.empty(ops.mdta_p1_blob_bytes(C), dtype=torch.uint8, device="cuda")
    ops.mdta_p1_pack(sd["b.attn.qkv.weight"].cuda(), sd["b.attn.qkv_dwconv.weight"].cuda(), blob)
    lnargs = None
    if ln:
        lnargs = (ops.ln_stats(xd), sd["b.norm1.body.weight"].cuda(), sd["b.norm1.body.bias"].cuda())
    for save in (False, True):
        G = torch.zeros(B, heads, c, c, device="cuda")
        ss = torch.zeros(B, 2 * C, device="cuda")
        v, pre, qkv = ops.mdta_p1(xd, blob, heads, G, ss, ln=lnargs, save=save)
        _close("v", v, qkv64[:, 2 * C:])
        _close("G", G, G64)
        _close("sumsq", ss, ss64)
        if save:
            _close("pre", pre, pre64)
            _close("qkv", qkv, qkv64)


@pytest.mark.parametrize("C,heads,H,W", [(96, 2, 32, 32), (48, 1, 16, 48), (96, 4, 24, 16)])
def test_mdta_forward_fused_matches_unfused_and_oracle(cuda_lib, C, heads, H, W):
    """engine.mdta_fwd takes the one-kernel phase 1 when nothing is kept for the backward; its block output must agree
    with the three-launch path on the same weights and with the oracle's MDTA (Net_Restormer.py:29-50)."""
    from oracle import restormer_ref as R
    from rcot_b200 import engine
    g = torch.Generator().manual_seed(7 + C + heads)
    sd = _params(C, heads, g)
    x = torch.randn(2, C, H, W, generator=g)
    sd64 = {k: v.double() for k, v in sd.items()}
    x64 = x.double()
    y64 = x64 + R.mdta(R.layer_norm_c(x64, sd64["b.norm1.body.weight"], sd64["b.norm1.body.bias"]), sd64, "b.attn.", heads)
    outs = []
    for fused in (True, False):
        ps = engine.ParamSet(dict(sd), "cuda")
        bs = engine.BlockSpec(ps, "b.", C, heads, has_ffn=False)
        if fused:
            ps.add_mdta("b.", C)
        else:
            ps.mdta.clear()
        ps.finalize()
        y, _ = engine.mdta_fwd(bs, x.cuda(), "b.norm1", True, False)
        outs.append(y)
        _close("y vs oracle", y, y64)
        if fused:                                     # "always" mode: the kept tensors feed the unfused backward
            engine.FUSED_MDTA_ALWAYS = True
            try:
                y2, ctx = engine.mdta_fwd(bs, x.cuda(), "b.norm1", True, True)
            finally:
                engine.FUSED_MDTA_ALWAYS = False
            _close("y (keep)", y2, y64)
            dx = engine.mdta_bwd(bs, x.cuda(), torch.ones_like(y2), "b.norm1", True, ctx)
            outs.append(dx)
        else:
            y3, ctx = engine.mdta_fwd(bs, x.cuda(), "b.norm1", True, True)
            outs.append(engine.mdta_bwd(bs, x.cuda(), torch.ones_like(y3), "b.norm1", True, ctx))
    torch.testing.assert_close(outs[0], outs[2], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(outs[1], outs[3], rtol=2e-3, atol=2e-4)
