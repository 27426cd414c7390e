"""One-kernel GDFN forward (csrc/gdfn_fused.cu; Net_Restormer.py:80-85 + norm2 + residual) against the fp64 oracle:
every supported (C, H, W) class including ragged tile counts and the benchmarked 128x128 / 64x64 shapes, the saved
hidden tensors u / g, the LayerNorm statistics it leaves for the next block, and the no-LN / no-residual form behind a
stand-alone FeedForward module."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(C, g):
    hid = int(C * 2.66)
    r = lambda *s: torch.randn(*s, generator=g)
    return {"b.norm2.body.weight": 1 + 0.2 * r(C), "b.norm2.body.bias": 0.2 * r(C),
            "b.ffn.project_in.weight": r(2 * hid, C, 1, 1) / C ** 0.5,
            "b.ffn.dwconv.weight": r(2 * hid, 1, 3, 3) / 3,
            "b.ffn.project_out.weight": r(C, hid, 1, 1) / hid ** 0.5}, hid


def _close(name, got, ref, rtol=1e-3, atol=1e-4):
    got = got.detach().cpu().double()
    err = (got - ref).abs()
    tol = atol * max(1.0, ref.abs().max().item()) + rtol * ref.abs()
    bad = (err > tol).sum().item()
    print(f"{name:10s} max_err={err.max().item():.3e} scale={ref.abs().max().item():.3e} bad={bad}")
    assert bad == 0, name


@pytest.mark.parametrize("C,B,H,W", [(48, 2, 8, 16), (96, 1, 16, 32), (96, 3, 24, 16), (48, 2, 128, 128),
                                     (96, 2, 128, 128), (96, 2, 64, 64), (96, 5, 40, 48)])
@pytest.mark.parametrize("ln,residual", [(True, True), (False, False)])
def test_gdfn_fused_forward(cuda_lib, C, B, H, W, ln, residual):
    from oracle import restormer_ref as R
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(C + H + W)
    sd, hid = _params(C, g)
    x = torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3
    x64 = x.double()
    sd64 = {k: v.double() for k, v in sd.items()}
    z64 = R.layer_norm_c(x64, sd64["b.norm2.body.weight"], sd64["b.norm2.body.bias"]) if ln else x64
    u64 = torch.nn.functional.conv2d(z64, sd64["b.ffn.project_in.weight"])
    d64 = torch.nn.functional.conv2d(u64, sd64["b.ffn.dwconv.weight"], padding=1, groups=2 * hid)
    g64 = R.gelu_exact(d64[:, :hid]) * d64[:, hid:]
    y64 = torch.nn.functional.conv2d(g64, sd64["b.ffn.project_out.weight"]) + (x64 if residual else 0)
    assert ops.gdfn_supported(C, H, W)
    xd = x.cuda()
    blob = torch.empty(ops.gdfn_blob_bytes(C, hid), dtype=torch.uint8, device="cuda")
    ops.gdfn_pack(sd["b.ffn.project_in.weight"].cuda(), sd["b.ffn.dwconv.weight"].cuda(),
                  sd["b.ffn.project_out.weight"].cuda(), blob)
    lnargs = None
    if ln:
        lnargs = (ops.ln_stats(xd), sd["b.norm2.body.weight"].cuda(), sd["b.norm2.body.bias"].cuda())
    for save in (False, True):
        y, u, gg = ops.gdfn_fwd(xd, blob, hid, ln=lnargs, residual=residual, stats_out=True, save=save)
        _close("y", y, y64)
        st = y._rcot_ln_stats.cpu().double().view(B, H, W, 2)
        _close("mean", st[..., 0], y64.mean(1), rtol=1e-4, atol=1e-4)
        _close("rstd", st[..., 1], 1.0 / torch.sqrt(y64.var(1, unbiased=False) + 1e-5), rtol=1e-3, atol=1e-4)
        if save:
            _close("u", u, u64)
            _close("g", gg, g64)


def test_gdfn_fused_matches_unfused_block_path(cuda_lib):
    """engine.gdfn_fwd takes the fused kernel by default; its output, kept tensors and the backward fed by them must agree
    with the three-launch path (RCOT_FUSED_GDFN=0 semantics) on the same weights."""
    from rcot_b200 import engine
    g = torch.Generator().manual_seed(3)
    C = 96
    sd, hid = _params(C, g)
    x = torch.randn(2, C, 32, 32, generator=g).cuda()
    dy = torch.randn(2, C, 32, 32, generator=g).cuda()
    outs = []
    for fused in (True, False):
        ps = engine.ParamSet({k: v for k, v in sd.items()}, "cuda")
        bs = engine.BlockSpec(ps, "b.", C, 1, has_attn=False)
        if fused:
            ps.add_gdfn("b.", C, hid)          # what BlockSpec does under RCOT_FUSED_GDFN=1
        else:
            ps.gdfn.clear()
        ps.finalize()
        engine.FUSED_GDFN_ALWAYS = fused           # the default ("auto") takes the fused kernel only when nothing is kept
        y, kept = engine.gdfn_fwd(bs, x, "b.norm2", True, keep=True)
        dx = engine.gdfn_bwd(bs, x, dy.clone(), "b.norm2", True, kept=kept)
        outs.append((y, dx, ps.grad.clone()))
    engine.FUSED_GDFN_ALWAYS = False
    for a, b in zip(outs[0], outs[1]):
        torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-4)


def test_inference_forward_uses_fused_kernel_and_matches(cuda_lib):
    """Default policy: a forward that keeps nothing (inference) runs GDFN as one kernel; same result as the three launches."""
    from rcot_b200 import engine, ops
    g = torch.Generator().manual_seed(5)
    C = 48
    sd, hid = _params(C, g)
    x = torch.randn(2, C, 64, 64, generator=g).cuda()
    ys, used = [], []
    for fused in (True, False):
        ps = engine.ParamSet({k: v for k, v in sd.items()}, "cuda")
        bs = engine.BlockSpec(ps, "b.", C, 1, has_attn=False)
        if not fused:
            ps.gdfn.clear()
        ps.finalize()
        ops.PROF = ops.Profiler()
        ys.append(engine.gdfn_fwd(bs, x, "b.norm2", True, keep=False))
        used.append("gdfn_fwd" in ops.PROF.summary())
        ops.PROF = None
    assert used == [True, False]
    torch.testing.assert_close(ys[0], ys[1], rtol=2e-4, atol=2e-4)
