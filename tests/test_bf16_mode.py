"""bf16-storage mode of the hidden tensors (SURVEY 8 f4 / BASELINE config c3): kernels reading / writing bf16 tensors
against the fp32 kernels on the same (bf16-representable) values, then a block and a full 128x128 iteration against the
oracle at the mode's STATED tolerance (network output rtol 2e-2 / atol 2e-3, Loss_mse 2e-2, Loss_T 5e-2 -- it carries the
potential's sign-like RMSprop steps --, gradients 5e-2 rel-L2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def bf16_mode():
    import rcot_b200
    rcot_b200.set_hidden_dtype("bf16")
    yield
    rcot_b200.set_hidden_dtype("fp32")


def _r(*s, g):
    return torch.randn(*s, generator=g)


def test_dwconv_kernels_bf16_storage(cuda_lib):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, hid, H, W = 2, 24, 32, 64
    u = _r(B, 2 * hid, H, W, g=g).cuda().bfloat16()
    w = (_r(2 * hid, 1, 3, 3, g=g) / 3).cuda()
    dg = _r(B, hid, H, W, g=g).cuda().bfloat16()
    uf, dgf = u.float(), dg.float()
    # plain (+ sums of squares of the STORED values), gate forward, gate backward, fused backward
    sq, sqf = torch.zeros(B, 2 * hid, device="cuda"), torch.zeros(B, 2 * hid, device="cuda")
    o = ops.dwconv(u, w, sumsq=sq, nsq=2 * hid)
    of = ops.dwconv(uf, w, sumsq=sqf, nsq=2 * hid)
    assert o.dtype == torch.bfloat16
    torch.testing.assert_close(o.float(), of.bfloat16().float(), rtol=0, atol=0)
    torch.testing.assert_close(sq, (o.float() ** 2).sum((2, 3)), rtol=1e-4, atol=1e-3)
    gt = ops.dwconv(u, w, mode=1)
    gtf = ops.dwconv(uf, w, mode=1)
    torch.testing.assert_close(gt.float(), gtf.bfloat16().float(), rtol=0, atol=0)
    gk, gkf = torch.empty_like(gt), torch.empty_like(gtf)
    dab = ops.dwconv(u, w, mode=2, dg=dg, g_out=gk, out=torch.empty_like(u))
    dabf = ops.dwconv(uf, w, mode=2, dg=dgf, g_out=gkf, out=torch.empty_like(uf))
    torch.testing.assert_close(dab.float(), dabf.bfloat16().float(), rtol=0, atol=0)
    torch.testing.assert_close(gk.float(), gkf.bfloat16().float(), rtol=0, atol=0)
    dw, dwf = torch.zeros(2 * hid, 1, 3, 3, device="cuda"), torch.zeros(2 * hid, 1, 3, 3, device="cuda")
    du = ops.dwconv_bwd(u, dab, w, dw)
    duf = ops.dwconv_bwd(uf, dab.float(), w, dwf)
    torch.testing.assert_close(du.float(), duf.bfloat16().float(), rtol=0, atol=0)
    torch.testing.assert_close(dw, dwf, rtol=1e-4, atol=1e-3)


def test_gemm_kernels_bf16_storage(cuda_lib):
    from rcot_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, C, N, H, W = 2, 96, 255, 32, 32
    x = _r(B, C, H, W, g=g).cuda()
    xb = x.bfloat16()
    wt = (_r(N, C, 1, 1, g=g) / C ** 0.5).cuda()
    pk = ops.pack_single(wt, "fwd")
    ref = ops.pm_gemm(xb.float(), pk.ptr(0), N)                               # fp32 kernel on the same values
    y1 = ops.pm_gemm(xb, pk.ptr(0), N)                                        # bf16 in, fp32 out
    torch.testing.assert_close(y1, ref, rtol=1e-4, atol=1e-4)
    y2 = ops.pm_gemm(xb, pk.ptr(0), N, out_dtype=torch.bfloat16)              # bf16 in, bf16 out
    torch.testing.assert_close(y2.float(), ref.bfloat16().float(), rtol=1e-2, atol=1e-2)
    stats = ops.ln_stats(x)
    gam, bet = (1 + 0.1 * _r(C, g=g)).cuda(), (0.1 * _r(C, g=g)).cuda()
    y3 = ops.pm_gemm(x, pk.ptr(0), N, ln=(stats, gam, bet), out_dtype=torch.bfloat16)   # fp32 in + LN, bf16 out
    ref3 = ops.pm_gemm(x, pk.ptr(0), N, ln=(stats, gam, bet))
    torch.testing.assert_close(y3.float(), ref3.bfloat16().float(), rtol=1e-2, atol=1e-2)
    # residual epilogue with a bf16 gather source (g -> y of GDFN): fp32 output
    res = _r(B, N, H, W, g=g).cuda()
    y4 = ops.pm_gemm(xb, pk.ptr(0), N, residual=res, stats_out=True)
    torch.testing.assert_close(y4, ref + res, rtol=1e-4, atol=1e-4)
    # pixel-as-K products: Gram (both bf16, per image, groups), dy v^T (b bf16), dW with LayerNorm (a bf16)
    q = _r(B, 96, H, W, g=g).cuda().bfloat16()
    k = _r(B, 96, H, W, g=g).cuda().bfloat16()
    G, Gf = torch.zeros(B, 2, 48, 48, device="cuda"), torch.zeros(B, 2, 48, 48, device="cuda")
    ops.pk_gemm(q, k, G, ldo=48, per_image=True, groups=2, out_gs=48 * 48)
    ops.pk_gemm(q.float(), k.float(), Gf, ldo=48, per_image=True, groups=2, out_gs=48 * 48)
    torch.testing.assert_close(G, Gf, rtol=1e-4, atol=1e-3)
    dy = _r(B, 96, H, W, g=g).cuda()
    P, Pf = torch.zeros(B, 96, 96, device="cuda"), torch.zeros(B, 96, 96, device="cuda")
    ops.pk_gemm(dy, k, P, ldo=96, per_image=True)
    ops.pk_gemm(dy, k.float(), Pf, ldo=96, per_image=True)
    torch.testing.assert_close(P, Pf, rtol=1e-4, atol=1e-3)
    gg = _r(B, 255, H, W, g=g).cuda().bfloat16()                              # dW_o = dy g^T, ldo = 255 -> operand swap
    dWo, dWof = torch.zeros(96, 255, device="cuda"), torch.zeros(96, 255, device="cuda")
    ops.pk_gemm(dy, gg, dWo, ldo=255)
    ops.pk_gemm(dy, gg.float(), dWof, ldo=255)
    torch.testing.assert_close(dWo, dWof, rtol=1e-4, atol=1e-3)
    du = _r(B, 510, H, W, g=g).cuda().bfloat16()
    dWi, dWif = torch.zeros(510, 96, device="cuda"), torch.zeros(510, 96, device="cuda")
    ops.pk_gemm(du, x, dWi, ldo=96, ln=(stats, gam, bet))
    ops.pk_gemm(du.float(), x, dWif, ldo=96, ln=(stats, gam, bet))
    torch.testing.assert_close(dWi, dWif, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("C,heads,H,W", [(48, 1, 32, 32), (96, 2, 64, 64), (96, 4, 16, 32), (192, 4, 16, 16)])
def test_block_bf16_mode(cuda_lib, bf16_mode, C, heads, H, W):
    """Block forward and backward with bf16 hidden tensors vs the fp64 oracle at the mode's tolerance
    (C = 192 keeps fp32 hidden tensors: it must still meet the fp32 tolerance)."""
    from oracle import restormer_ref as R
    from rcot_b200 import engine
    from tests.test_block import _block_params
    g = torch.Generator().manual_seed(C + heads)
    sd = _block_params(C, heads, g)
    x = torch.randn(2, C, H, W, generator=g)
    dy = torch.randn(2, C, H, W, generator=g)
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    y64 = R.transformer_block(x64, sd64, "b.", heads)
    y64.backward(dy.double())
    for save in (True, False):
        ps = engine.ParamSet(dict(sd), "cuda")
        bs = engine.BlockSpec(ps, "b.", C, heads)
        ps.finalize()
        tape = engine.Tape(save_hidden=save)
        xd = x.cuda()
        y = engine.block_fwd(bs, xd, tape)
        leaves = tape.backward(y, dy.cuda().clone())
        tol = 2e-2 if C <= 96 else 2e-3

        def rel(a, b):
            return ((a.detach().cpu().double().reshape(b.shape) - b).norm() / b.norm()).item()
        errs = {"y": rel(y, y64.detach()), "dx": rel(tape.grad_of(leaves, xd), x64.grad)}
        for k in sd:
            errs[k] = rel(ps.g[k], sd64[k].grad)
        worst = max(errs, key=errs.get)
        print(f"C={C} save={save}: y {errs['y']:.2e} dx {errs['dx']:.2e} worst {worst} {errs[worst]:.2e}")
        assert errs["y"] < tol / 4 and errs["dx"] < tol, errs
        for k, e in errs.items():
            assert e < (5 * tol if "temperature" in k else 2.5 * tol), (k, e)


def test_full_iteration_bf16_mode(cuda_lib, bf16_mode):
    from oracle import restormer_ref as R
    from oracle import train_ref
    from rcot_b200.engine import Tape
    from rcot_b200.train_step import OTTrainStep
    from tests.test_bench_size import _batch, _nets
    P, B = 128, 2
    Tp, Fp, T_sd, F_sd = _nets(P)
    deg, tgt = _batch(11, B, P)
    de_id, alpha = torch.tensor([1, 4]), torch.tensor([0.25, 0.7])
    step = OTTrainStep(Tp, Fp, "RMSprop")
    step.capture = {}
    Tflat0 = Tp.ps.flat.clone()
    r = step.iteration(deg.cuda(), tgt.cuda(), de_id.cuda(), alpha.cuda(), False, 1e-4)
    o = train_ref.train_iteration(T_sd, F_sd, {}, {}, deg, tgt, de_id, alpha, 1e-4, 1.0, 10000.0, False)
    torch.testing.assert_close(r["out"].cpu(), o["out"], rtol=2e-2, atol=2e-3)
    assert abs(r["loss_mse"].item() - o["loss_mse"]) <= 2e-2 * abs(o["loss_mse"]), (r["loss_mse"].item(), o["loss_mse"])
    # Loss_T = -mean f(T(x)) + cost, with f the potential AFTER its two sign-like RMSprop steps (each weight moves by
    # +-10*lr whatever |g|; tests/test_bench_size.py explains the +-0.3 % this carries in fp32 mode). With bf16 hidden
    # tensors T(x) is perturbed ~100x more (1e-3 instead of 1e-5 relative), more near-zero gradients of the potential
    # flip sign, and -- because an fp32-level change of summation order (split-K atomics) can move a stored value to the
    # neighbouring bf16 -- the SAME binary on the SAME inputs spreads: seven runs on a B200 gave 624.6 ... 644.8 against
    # the oracle's 638.3 (-2.1 % ... +1.0 %, sigma 1.2 %; scripts/gpu/diag_bf16_spread.sh). 5e-2 is 4 sigma of that spread.
    assert abs(r["loss_T"].item() - o["loss_T"]) <= 5e-2 * abs(o["loss_T"]), (r["loss_T"].item(), o["loss_T"])
    # Flat T gradient.  Inside the iteration dL/dout carries dF/dout of the potential after its two sign-like steps (see
    # above), so -- like tests/test_bench_size.py -- the mode's 5e-2 is asserted on the exact half: the bf16-mode T
    # forward + backward from the initial weights, driven by the ORACLE's dL/dout.  The in-iteration buffer (below 5e-2
    # in the runs that printed it) only gets a gross-error bound: a flaky -x stop is worth less than that assert.
    def flat_err(flat):
        num = den = 0.0
        for k, off in Tp.ps.offsets.items():
            ref = o["grads_T"].get(k)
            if ref is None:
                continue
            got = flat[off:off + ref.numel()].view(ref.shape).cpu().double()
            num += (got - ref.double()).pow(2).sum().item()
            den += ref.double().pow(2).sum().item()
        return (num / den) ** 0.5
    err_in = flat_err(step.capture["T"])
    out_o = o["out"].clone().requires_grad_(True)
    with torch.enable_grad():
        lossT, _ = R.transport_loss(out_o, deg, tgt, R.fnet_forward(F_sd, out_o), de_id, 1.0, 10000.0, False)
    lossT.backward()
    Tp.ps.flat.copy_(Tflat0)
    Tp.ps.repack()
    Tp.ps.zero_grad()
    tape = Tape(save_hidden=True)
    out = Tp.forward(deg.cuda(), tape)
    tape.backward(out, out_o.grad.cuda())
    err = flat_err(Tp.ps.grad)
    print(f"bf16 mode: flat T-gradient rel-L2 {err:.3e} (oracle dL/dout), {err_in:.3e} (inside the iteration); "
          f"out max err {(r['out'].cpu() - o['out']).abs().max().item():.2e}")
    assert err < 5e-2, err
    assert err_in < 0.5, err_in
