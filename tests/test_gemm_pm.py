"""Pixel-as-M tcgen05 GEMM (rcot_pm_gemm) against fp64 F.conv2d: 1x1 (+LN, +concat, +residual),
dense 3x3 / 4x4 s2 / 5x5 forward and data-gradient geometry, N > 256 and ragged K."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol=1e-3, atol=1e-4):
    got = got.detach().cpu().double()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    print(f"max_abs_err={err.max().item():.3e} ref_absmax={ref.abs().max().item():.3e} bad={bad}/{err.numel()}")
    assert bad == 0


@pytest.mark.parametrize("Cin,Cout,H,W", [(48, 144, 16, 16), (96, 510, 8, 24), (255, 96, 16, 8), (384, 2042, 8, 8),
                                          (1021, 384, 4, 4), (3, 48, 5, 7)])
def test_pointwise(cuda_lib, Cin, Cout, H, W):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    x = torch.randn(2, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5
    ref = F.conv2d(x.double(), w.double())
    wd = w.cuda()
    pk = ops.pack_single(wd, "fwd")
    out = ops.pm_gemm(x.cuda(), pk.ptr(0), Cout)
    _close(out, ref)
    # transposed (data gradient of the 1x1): dx = W^T dy
    dy = torch.randn(2, Cout, H, W, generator=g)
    refdx = F.conv_transpose2d(dy.double(), w.double())
    pkT = ops.pack_single(wd, "dgrad")
    dx = ops.pm_gemm(dy.cuda(), pkT.ptr(0), Cin)
    _close(dx, refdx)


def test_pointwise_ln_cat_residual(cuda_lib):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, C1, C2, H, W, N = 2, 96, 192, 12, 20, 192
    x1 = torch.randn(B, C1, H, W, generator=g)
    x2 = torch.randn(B, C2, H, W, generator=g)
    w = torch.randn(N, C1 + C2, 1, 1, generator=g) / (C1 + C2) ** 0.5
    res = torch.randn(B, N, H, W, generator=g)
    ref = F.conv2d(torch.cat([x1, x2], 1).double(), w.double()) + res.double()
    pk = ops.pack_single(w.cuda(), "fwd")
    out = ops.pm_gemm(x1.cuda(), pk.ptr(0), N, x2=x2.cuda(), residual=res.cuda())
    _close(out, ref)
    # LayerNorm prologue on a single input, written into a channel slice of a wider tensor
    gamma = torch.randn(C1, generator=g)
    beta = torch.randn(C1, generator=g)
    w1 = torch.randn(N, C1, 1, 1, generator=g) / C1 ** 0.5
    xd = x1.double()
    mu = xd.mean(1, keepdim=True)
    var = ((xd - mu) ** 2).mean(1, keepdim=True)
    rstd = 1 / torch.sqrt(var + 1e-5)
    z = (xd - mu) * rstd * gamma.double().view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1)
    ref2 = F.conv2d(z, w1.double())
    stats = torch.stack([mu.flatten(1), rstd.flatten(1)], -1).float().contiguous().cuda()
    pk1 = ops.pack_single(w1.cuda(), "fwd")
    wide = torch.zeros(B, N + 16, H, W, device="cuda")
    ops.pm_gemm(x1.cuda(), pk1.ptr(0), N, ln=(stats, gamma.cuda(), beta.cuda()), out=wide, out_coff=16)
    _close(wide[:, 16:], ref2)
    assert wide[:, :16].abs().max().item() == 0


@pytest.mark.parametrize("Cin,Cout,k,s,p,H,W,bias,act", [
    (48, 24, 3, 1, 1, 16, 16, False, False), (3, 64, 5, 1, 2, 16, 16, True, True),
    (64, 64, 4, 2, 1, 16, 16, True, True), (256, 512, 3, 1, 1, 8, 8, False, True),
    (512, 512, 4, 2, 1, 4, 4, False, True), (96, 3, 3, 1, 1, 10, 14, False, False),
    (48, 48, 3, 1, 1, 12, 8, False, False), (16, 80, 5, 1, 2, 8, 8, True, True)])
def test_conv_fwd_dgrad(cuda_lib, Cin, Cout, k, s, p, H, W, bias, act):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(Cin + Cout + k)
    x = torch.randn(2, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) if bias else None
    ref = F.conv2d(x.double(), w.double(), None if b is None else b.double(), stride=s, padding=p)
    if act:
        ref = F.leaky_relu(ref, 0.2)
    wd = w.cuda()
    pk = ops.pack_single(wd, "fwd")
    out = ops.pm_gemm(x.cuda(), pk.ptr(0), Cout, ks=k, stride=s, pad=p, bias=None if b is None else b.cuda(), act=act)
    _close(out, ref)
    dy = torch.randn_like(ref).float()
    refdx = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), stride=s, padding=p)
    pkT = ops.pack_single(wd, "dgrad")
    dx = ops.pm_gemm(dy.cuda(), pkT.ptr(0), Cin, ks=k, stride=s, pad=p, mode=1, out_hw=(H, W))
    _close(dx, refdx)
    # tap-major K order (ky, kx, channel): the fast producer path used whenever channels % 16 == 0
    if Cin % 16 == 0:
        pkt = ops.pack_single(wd, "fwd_tap")
        out_t = ops.pm_gemm(x.cuda(), pkt.ptr(0), Cout, ks=k, stride=s, pad=p, bias=None if b is None else b.cuda(),
                            act=act, tap_major=True)
        _close(out_t, ref)
    if Cout % 16 == 0:
        pktT = ops.pack_single(wd, "dgrad_tap")
        dx_t = ops.pm_gemm(dy.cuda(), pktT.ptr(0), Cin, ks=k, stride=s, pad=p, mode=1, out_hw=(H, W), tap_major=True)
        _close(dx_t, refdx)
    # sign-mask epilogue + accumulate: out = prev + dx * (mask>0 ? 1 : 0.2)
    mask = torch.randn(2, Cin, H, W, generator=g)
    prev = torch.randn(2, Cin, H, W, generator=g)
    acc = prev.cuda().clone()
    ops.pm_gemm(dy.cuda(), pkT.ptr(0), Cin, ks=k, stride=s, pad=p, mode=1, out_hw=(H, W), out=acc,
                mask_y=mask.cuda(), accumulate=True)
    _close(acc, prev.double() + refdx * torch.where(mask > 0, 1.0, 0.2).double())
    # mask alone: the prefetching epilogue of F_net's data gradients (and the parity-class tiles when k=4, s=2)
    for tap in ([False, True] if Cout % 16 == 0 else [False]):
        pkm = ops.pack_single(wd, "dgrad_tap" if tap else "dgrad")
        only = ops.pm_gemm(dy.cuda(), pkm.ptr(0), Cin, ks=k, stride=s, pad=p, mode=1, out_hw=(H, W), mask_y=mask.cuda(),
                           tap_major=tap)
        _close(only, refdx * torch.where(mask > 0, 1.0, 0.2).double())


def test_bf16_single_term(cuda_lib):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 96, 16, 16, generator=g)
    w = torch.randn(192, 96, 1, 1, generator=g) / 96 ** 0.5
    ref = F.conv2d(x.double(), w.double())
    pk = ops.pack_single(w.cuda(), "fwd")
    out = ops.pm_gemm(x.cuda(), pk.ptr(0), 192, terms=1)
    _close(out, ref, rtol=2e-2, atol=2e-2)  # stated bf16-compute tolerance


@pytest.mark.parametrize("C1,C2,N,H,W,ln", [(96, 0, 510, 16, 24, True), (64, 32, 192, 8, 16, False), (48, 0, 254, 16, 16, True),
                                            (255, 0, 96, 16, 8, False), (127, 0, 48, 8, 32, False), (192, 0, 192, 16, 16, True)])
def test_pointwise_tma_staged(cuda_lib, C1, C2, N, H, W, ln):
    """1x1 GEMMs on feature maps with H*W % 128 == 0 take the TMA-staged producer path (bulk copies of the raw fp32
    tile into a shared-memory ring): ragged K (48, 127, 255), concat, LayerNorm prologue, residual, several tiles per
    image and several images per launch, against fp64."""
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(C1 + C2 + N)
    B = 3
    x1 = torch.randn(B, C1, H, W, generator=g) + 0.3
    x2 = torch.randn(B, C2, H, W, generator=g) if C2 else None
    w = torch.randn(N, C1 + C2, 1, 1, generator=g) / (C1 + C2) ** 0.5
    res = torch.randn(B, N, H, W, generator=g)
    xin = x1.double() if x2 is None else torch.cat([x1, x2], 1).double()
    kw = {}
    if ln:
        gamma, beta = torch.randn(C1, generator=g), torch.randn(C1, generator=g)
        mu = xin.mean(1, keepdim=True)
        rstd = 1 / torch.sqrt(((xin - mu) ** 2).mean(1, keepdim=True) + 1e-5)
        xin = (xin - mu) * rstd * gamma.double().view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1)
        kw["ln"] = (ops.ln_stats(x1.cuda()), gamma.cuda(), beta.cuda())
    ref = F.conv2d(xin, w.double()) + res.double()
    pk = ops.pack_single(w.cuda(), "fwd")
    out = ops.pm_gemm(x1.cuda(), pk.ptr(0), N, x2=None if x2 is None else x2.cuda(), residual=res.cuda(), **kw)
    _close(out, ref)
    # channel-slice input view (as MDTA's v = qkv[:, 2C:]): base pointer offset inside a wider tensor
    wide = torch.randn(B, C1 + 16, H, W, generator=g).cuda()
    sl = wide[:, 16:]
    ws = torch.randn(N, C1, 1, 1, generator=g) / C1 ** 0.5
    pk2 = ops.pack_single(ws.cuda(), "fwd")
    _close(ops.pm_gemm(sl, pk2.ptr(0), N), F.conv2d(sl.cpu().double(), ws.double()))
