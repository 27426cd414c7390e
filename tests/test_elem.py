"""Unit parity of the building-block kernels against fp64 PyTorch: pixel-as-K GEMM (weight
gradients, per-image Grams), LayerNorm stats/backward, depthwise 3x3 variants, pixel shuffle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(name, got, ref, rtol=1e-3, atol=1e-4):
    got = got.detach().cpu().double().reshape(ref.shape)
    scale = max(1.0, ref.abs().max().item())
    err = (got - ref).abs()
    bad = (err > atol * scale + rtol * ref.abs()).sum().item()
    print(f"{name:24s} max_err={err.max().item():.3e} scale={scale:.3e} bad={bad}/{err.numel()}")
    assert bad == 0, name


@pytest.mark.parametrize("Cin,Cout,k,s,p,H,W", [(48, 24, 3, 1, 1, 16, 16), (3, 64, 5, 1, 2, 16, 16),
                                                (64, 64, 4, 2, 1, 16, 16), (256, 512, 3, 1, 1, 8, 8),
                                                (512, 512, 4, 2, 1, 4, 4), (512, 512, 4, 2, 1, 2, 2),
                                                (96, 3, 3, 1, 1, 10, 14), (96, 300, 1, 1, 0, 8, 8),
                                                # ldo % 4 != 0: operands swapped, transposed (coalesced) reductions
                                                (127, 48, 1, 1, 0, 16, 16), (255, 96, 1, 1, 0, 8, 16),
                                                (510, 192, 1, 1, 0, 8, 8)])
def test_conv_wgrad(cuda_lib, Cin, Cout, k, s, p, H, W):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(Cin + Cout + k)
    B = 3
    x = torch.randn(B, Cin, H, W, generator=g)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    dy = torch.randn(B, Cout, Ho, Wo, generator=g)
    ref = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, k, k), dy.double(), stride=s, padding=p)
    prev = torch.randn(Cout, Cin, k, k, generator=g)
    out = prev.cuda().clone()
    ops.pk_gemm(dy.cuda(), x.cuda(), out, ldo=Cin * k * k, ks=k, stride=s, pad=p)
    _close("dW", out, prev.double() + ref)


def test_wgrad_ln_and_concat(cuda_lib):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, C, H, W, N = 2, 96, 8, 16, 510
    x = torch.randn(B, C, H, W, generator=g)
    du = torch.randn(B, N, H, W, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    xd = x.double()
    mu = xd.mean(1, keepdim=True)
    rstd = 1 / torch.sqrt(((xd - mu) ** 2).mean(1, keepdim=True) + 1e-5)
    z = (xd - mu) * rstd * gamma.double().view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1)
    ref = torch.einsum("bnhw,bchw->nc", du.double(), z)
    stats = ops.ln_stats(x.cuda())
    _close("stats", stats, torch.stack([mu.flatten(1), rstd.flatten(1)], -1))
    out = torch.zeros(N, C, device="cuda")
    ops.pk_gemm(du.cuda(), x.cuda(), out, ldo=C, ln=(stats, gamma.cuda(), beta.cuda()))
    _close("dW_ln", out, ref)
    x2 = torch.randn(B, 48, H, W, generator=g)
    ref2 = torch.einsum("bnhw,bchw->nc", du.double(), torch.cat([x, x2], 1).double())
    out2 = torch.zeros(N, C + 48, device="cuda")
    ops.pk_gemm(du.cuda(), x.cuda(), out2, ldo=C + 48, b2=x2.cuda())
    _close("dW_cat", out2, ref2)


@pytest.mark.parametrize("C,heads,HW", [(48, 1, (16, 16)), (96, 4, (8, 8)), (384, 8, (4, 4)), (384, 4, (4, 8)),
                                           # H*W % 32 == 0 and <= 128 B-operand rows: the TMA-staged kernel
                                           (96, 1, (16, 24)), (192, 4, (8, 16)), (96, 2, (32, 32))])
def test_gram_per_image(cuda_lib, C, heads, HW):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(C)
    B = 2
    qkv = torch.randn(B, 3 * C, *HW, generator=g)
    c = C // heads
    q = qkv[:, :C].reshape(B, heads, c, -1).double()
    k = qkv[:, C:2 * C].reshape(B, heads, c, -1).double()
    ref = q @ k.transpose(-1, -2)
    d = qkv.cuda()
    G = torch.zeros(B, heads, c, c, device="cuda")
    ops.pk_gemm(d[:, :C], d[:, C:2 * C], G, ldo=c, per_image=True, groups=heads, out_gs=c * c)
    _close("gram", G, ref)
    v = qkv[:, 2 * C:].flatten(2).double()
    dy = torch.randn(B, C, *HW, generator=g)
    P = torch.zeros(B, C, C, device="cuda")
    ops.pk_gemm(dy.cuda(), d[:, 2 * C:], P, ldo=C, per_image=True)
    _close("P", P, dy.flatten(2).double() @ v.transpose(1, 2))


def test_layernorm_bwd(cuda_lib):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, C, H, W = 2, 96, 8, 12
    x = torch.randn(B, C, H, W, generator=g) * 2 + 0.5
    dz = torch.randn(B, C, H, W, generator=g)
    dy = torch.randn(B, C, H, W, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    x64 = x.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    from oracle import restormer_ref as R
    z = R.layer_norm_c(x64, g64, b64)
    z.backward(dz.double())
    stats = ops.ln_stats(x.cuda())
    dgam, dbet = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = ops.ln_bwd(dz.cuda(), x.cuda(), stats, gamma.cuda(), dgam, dbet, dy=dy.cuda())
    _close("dx", dx, x64.grad + dy.double())
    _close("dgamma", dgam, g64.grad)
    _close("dbeta", dbet, b64.grad)


def test_dwconv_variants(cuda_lib):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(6)
    B, hid, H, W = 2, 127, 10, 12
    Cn = 2 * hid
    u = torch.randn(B, Cn, H, W, generator=g)
    w = torch.randn(Cn, 1, 3, 3, generator=g) / 3
    u64 = u.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    full = F.conv2d(u64, w64, padding=1, groups=Cn)
    a, b = full.chunk(2, 1)
    gate = F.gelu(a) * b
    dgate = torch.randn(B, hid, H, W, generator=g)
    gate.backward(dgate.double())
    ud, wd = u.cuda(), w.cuda()
    sumsq = torch.zeros(B, 100, device="cuda")
    out = ops.dwconv(ud, wd, sumsq=sumsq, nsq=100)
    _close("dw plain", out, full.detach())
    _close("sumsq", sumsq, (full.detach()[:, :100] ** 2).sum((2, 3)))
    _close("gate", ops.dwconv(ud, wd, mode=1), gate.detach())
    gout = torch.empty(B, hid, H, W, device="cuda")
    dab = ops.dwconv(ud, wd, mode=2, dg=dgate.cuda(), g_out=gout, out=torch.empty_like(ud))
    _close("g_out", gout, gate.detach())
    du = ops.dwconv(dab, wd, flip=True)
    _close("du", du, u64.grad)
    dw = torch.zeros_like(wd)
    ops.dwconv_wgrad(ud, dab, dw)
    _close("dw_w", dw, w64.grad)


def test_pixel_shuffle_axpby(cuda_lib):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 24, 6, 10, generator=g)
    xd = x.cuda()
    assert torch.equal(ops.pixel_shuffle(xd).cpu(), F.pixel_shuffle(x, 2))
    assert torch.equal(ops.pixel_shuffle(xd, inverse=True).cpu(), F.pixel_unshuffle(x, 2))
    wide = torch.zeros(2, 10, 12, 20, device="cuda")
    ops.pixel_shuffle(xd, out=wide[:, 2:8])
    assert torch.equal(wide[:, 2:8].cpu(), F.pixel_shuffle(x, 2)) and wide[:, :2].abs().max() == 0
    y = torch.randn(2, 24, 6, 10, generator=g)
    torch.testing.assert_close(ops.axpby(xd, y.cuda(), 1.0, 0.8).cpu(), x + 0.8 * y)
    al = torch.rand(2, generator=g)
    torch.testing.assert_close(ops.axpby(xd, y.cuda(), a_vec=al.cuda()).cpu(),
                               al.view(2, 1, 1, 1) * x + (1 - al.view(2, 1, 1, 1)) * y)


@pytest.mark.parametrize("B,Cn,H,W", [(3, 37, 10, 16), (16, 600, 32, 32), (5, 9, 64, 64)])
def test_dwconv_bwd_fused(cuda_lib, B, Cn, H, W):
    """din and dW of the depthwise conv from one kernel vs autograd (small planes with several images per thread,
    large planes with several patches per thread)."""
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(16)
    x = torch.randn(B, Cn, H, W, generator=g)
    w = torch.randn(Cn, 1, 3, 3, generator=g) / 3
    dout = torch.randn(B, Cn, H, W, generator=g)
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    F.conv2d(x64, w64, padding=1, groups=Cn).backward(dout.double())
    prev = torch.randn(Cn, 1, 3, 3, generator=g)
    dw = prev.cuda().clone()
    din = ops.dwconv_bwd(x.cuda(), dout.cuda(), w.cuda(), dw)
    _close("din", din, x64.grad)
    _close("dw", dw, prev.double() + w64.grad)


def test_dwconv_odd_sizes(cuda_lib):
    """Whole-image inference reaches feature maps like 25x39: the scalar fallback path."""
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(21)
    B, hid, H, W = 2, 9, 5, 7
    u = torch.randn(B, 2 * hid, H, W, generator=g)
    w = torch.randn(2 * hid, 1, 3, 3, generator=g) / 3
    full = F.conv2d(u.double(), w.double(), padding=1, groups=2 * hid)
    a, b = full.chunk(2, 1)
    sumsq = torch.zeros(B, 6, device="cuda")
    _close("plain", ops.dwconv(u.cuda(), w.cuda(), sumsq=sumsq, nsq=6), full)
    _close("sumsq", sumsq, (full[:, :6] ** 2).sum((2, 3)))
    _close("gate", ops.dwconv(u.cuda(), w.cuda(), mode=1), F.gelu(a) * b)


@pytest.mark.parametrize("B,K,O", [(64, 8192, 2048), (32, 2048, 64), (5, 64, 1), (70, 136, 37), (3, 50, 7)])
def test_linear_kernels(cuda_lib, B, K, O):
    """F_net's fully connected tail (Net_Restormer.py:496-498,512-520): tiled fp32 GEMM forward (bias, LeakyReLU,
    mask; reduction split with atomics on the wide layer), data gradient and weight gradient vs fp64."""
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(B + K + O)
    x = torch.randn(B, K, generator=g)
    W = torch.randn(O, K, generator=g) / K ** 0.5
    bias = torch.randn(O, generator=g)
    dy = torch.randn(B, O, generator=g)
    mo = torch.randn(B, O, generator=g)
    mk = torch.randn(B, K, generator=g)
    x64, W64, b64 = x.double(), W.double(), bias.double()
    lin = x64 @ W64.t() + b64
    xd, Wd, bd = x.cuda(), W.cuda(), bias.cuda()
    _close("fwd bias", ops.linear_fwd(xd, Wd, bd), lin)
    _close("fwd nobias", ops.linear_fwd(xd, Wd), x64 @ W64.t())
    _close("fwd act", ops.linear_fwd(xd, Wd, bd, act=True), F.leaky_relu(lin, 0.2))
    fac_o = torch.where(mo > 0, 1.0, 0.2).double()
    _close("fwd mask", ops.linear_fwd(xd, Wd, mask=mo.cuda()), (x64 @ W64.t()) * fac_o)
    fac_k = torch.where(mk > 0, 1.0, 0.2).double()
    _close("dgrad", ops.linear_dgrad(dy.cuda(), Wd), dy.double() @ W64)
    _close("dgrad mask", ops.linear_dgrad(dy.cuda(), Wd, mask=mk.cuda()), (dy.double() @ W64) * fac_k)
    dW = torch.ones(O, K, device="cuda")
    db = torch.ones(O, device="cuda")
    ops.linear_wgrad(dy.cuda(), xd, dW, db)
    _close("wgrad", dW, 1 + dy.double().t() @ x64)
    _close("bgrad", db, 1 + dy.double().sum(0))


@pytest.mark.parametrize("B,hid,H,W,with_g", [(2, 5, 32, 32, True), (2, 3, 48, 64, False), (1, 4, 8, 32, True),
                                              (2, 2, 128, 128, True)])
def test_gdfn_mid_bwd_fused(cuda_lib, B, hid, H, W, with_g):
    """Gate backward + transposed depthwise conv + its weight gradient in one kernel vs autograd (fp64), and
    against the two-kernel form it replaces (same taps in the same order)."""
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(H + W + hid)
    Cn = 2 * hid
    u = torch.randn(B, Cn, H, W, generator=g)
    w = torch.randn(Cn, 1, 3, 3, generator=g) / 3
    dgate = torch.randn(B, hid, H, W, generator=g)
    u64, w64 = u.double().requires_grad_(True), w.double().requires_grad_(True)
    a, b = F.conv2d(u64, w64, padding=1, groups=Cn).chunk(2, 1)
    gate = F.gelu(a) * b
    gate.backward(dgate.double())
    ud, wd, dgd = u.cuda(), w.cuda(), dgate.cuda()
    assert ops.gdfn_mid_ok(ud)
    prev = torch.randn(Cn, 1, 3, 3, generator=g)
    dw = prev.cuda().clone()
    gout = torch.full((B, hid, H, W), float("nan"), device="cuda") if with_g else None
    du = ops.gdfn_mid_bwd(ud, dgd, wd, dw, g_out=gout)
    _close("du", du, u64.grad)
    _close("dw", dw, prev.double() + w64.grad, rtol=2e-3)
    if with_g:
        _close("g_out", gout, gate.detach())
    # the two-kernel form: same taps in the same order -> identical du and g
    g2 = torch.empty(B, hid, H, W, device="cuda")
    dab = ops.dwconv(ud, wd, mode=2, dg=dgd, g_out=g2, out=torch.empty_like(ud))
    dw2 = prev.cuda().clone()
    du2 = ops.dwconv_bwd(ud, dab, wd, dw2)
    torch.testing.assert_close(du, du2, rtol=1e-5, atol=1e-6)
    if with_g:
        torch.testing.assert_close(gout, g2, rtol=1e-5, atol=1e-6)
    _close("dw vs two-kernel", dw, dw2.cpu().double(), rtol=2e-3)


@pytest.mark.parametrize("B,CA,CB,H,W,off", [(3, 254, 48, 8, 8, 0), (2, 1020, 192, 8, 8, 0), (2, 300, 384, 4, 8, 0),
                                             (3, 288, 96, 8, 16, 1), (1, 144, 48, 16, 16, 0)])
def test_wgrad_ln_multi_m(cuda_lib, B, CA, CB, H, W, off):
    """The multi-M pixel-as-K kernel (MT = 2 / 4 accumulators per CTA, NBT = 1 / 2, two N tiles, chunk walker across
    images, vector and scalar reductions) against fp64: dW = dOut LN(x)^T as in the blocks' 1x1 weight gradients."""
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(CA + CB + off)
    x = torch.randn(B, CB, H, W, generator=g) + 0.5
    du = torch.randn(B, CA, H, W, generator=g)
    gamma, beta = torch.randn(CB, generator=g), torch.randn(CB, generator=g)
    xd = x.double()
    mu = xd.mean(1, keepdim=True)
    rstd = 1 / torch.sqrt(((xd - mu) ** 2).mean(1, keepdim=True) + 1e-5)
    z = (xd - mu) * rstd * gamma.double().view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1)
    ref = torch.einsum("bnhw,bchw->nc", du.double(), z)
    stats = ops.ln_stats(x.cuda())
    prev = torch.randn(CA * CB + off, generator=g)
    buf = prev.cuda().clone()
    out = buf[off:].view(CA, CB)          # off = 1: rows not 16-byte aligned -> scalar reductions
    ops.pk_gemm(du.cuda(), x.cuda(), out, ldo=CB, ln=(stats, gamma.cuda(), beta.cuda()))
    _close("dW_ln_mm", out, prev[off:].view(CA, CB).double() + ref)
