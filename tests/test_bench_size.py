"""Parity at the sizes bench.py runs (VERDICT r1 "parity gaps"): one FULL adversarial iteration at 128x128 against the
CPU oracle (paired and unpaired), and batch consistency at batch 32 -- the persistent multi-tile loops, N slices and the
hidden-tensor saving mode that only the benchmark exercised before."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nets(P):
    import Net_Restormer as N
    from rcot_b200.fnet import FnetProgram
    from rcot_b200.tnet import TnetProgram
    torch.manual_seed(0)
    T = N.T_net(decoder=True)
    F = N.F_net(patch_size=P)
    T_sd = {k: v.detach().clone() for k, v in T.state_dict().items()}
    F_sd = {k: v.detach().clone() for k, v in F.state_dict().items()}
    Tp = TnetProgram({k: v.detach().cuda() for k, v in T.named_parameters()}, "cuda")
    Fp = FnetProgram({k: v.detach().cuda() for k, v in F.named_parameters()}, "cuda", P)
    return Tp, Fp, T_sd, F_sd


def _batch(seed, B, P):
    g = torch.Generator().manual_seed(seed)
    tgt = torch.floor(torch.rand(B, 3, P, P, generator=g) * 255) / 255
    deg = torch.floor(torch.clamp(tgt * 255 + 25 * torch.randn(B, 3, P, P, generator=g), 0, 255)) / 255
    return deg, tgt


def _cmp_grads(ps, flat, ref, tol, what, tol_tensor=None):
    """Whole flat gradient: rel-L2 < tol.  Per tensor: rel-L2 < tol_tensor (default 10 * tol): single tensors of the
    potential's gradients are small differences of large, nearly cancelling real/fake contributions (first conv layer:
    |grad| ~ 1e-2 of either term), so their relative error is the products' 1e-4-class error amplified by that ratio."""
    tol_tensor = 10 * tol if tol_tensor is None else tol_tensor
    worst, worst_k, num, den, rows = 0.0, None, 0.0, 0.0, []
    for k, o in ps.offsets.items():
        r = ref.get(k)
        got = flat[o:o + ps.p[k].numel()].view(ps.p[k].shape).cpu().double()
        if r is None:
            assert got.abs().max().item() == 0, (what, k)
            continue
        r = r.double()
        d2, r2 = (got - r).pow(2).sum().item(), r.pow(2).sum().item()
        num, den = num + d2, den + r2
        err = d2 ** 0.5 / max(r2 ** 0.5, 1e-30)
        rows.append((err, k, r2 ** 0.5))
        if err > worst:
            worst, worst_k = err, k
    rows.sort(reverse=True)
    tot = (num / max(den, 1e-60)) ** 0.5
    print(f"{what}: flat rel-L2 {tot:.3e}; worst tensors: " + ", ".join(f"{k} {e:.2e} (|g|={n:.2e})" for e, k, n in rows[:4]))
    assert tot < tol, (what, tot)
    for e, k, n in rows:
        # attn.temperature: O(1e-3) sums of cancelling O(1) terms (fp32 noise of the oracle itself)
        assert e < tol_tensor or "temperature" in k or n * e < 1e-6, (what, k, e)


@pytest.mark.parametrize("paired", [True, False])
def test_full_iteration_128(cuda_lib, paired):
    """P=128, B=2, de_id = [1, 4] (both Fourier branches): printed losses (rtol 2e-4) and every gradient tensor of the
    three objectives (rel-L2 < 3e-3) vs oracle/train_ref.py.  The paired transport gradient is compared through the
    oracle's own dL/dout (sign(out - target) of the 10000 x L1 term flips on fp32 ties, see test_train_step.py)."""
    from oracle import restormer_ref as R
    from oracle import train_ref
    from rcot_b200.engine import Tape
    from rcot_b200.train_step import OTTrainStep
    P, B = 128, 2
    Tp, Fp, T_sd, F_sd = _nets(P)
    F0 = {k: v.clone() for k, v in F_sd.items()}
    deg, tgt = _batch(11, B, P)
    de_id = torch.tensor([1, 4])
    alpha = torch.tensor([0.25, 0.7])
    step = OTTrainStep(Tp, Fp, "RMSprop", sigma=1.0, Sigma=10000.0)
    step.capture = {}
    Tflat0, Fflat0 = Tp.ps.flat.clone(), Fp.ps.flat.clone()
    r = step.iteration(deg.cuda(), tgt.cuda(), de_id.cuda(), alpha.cuda(), paired, 1e-4)
    o = train_ref.train_iteration(T_sd, F_sd, {}, {}, deg, tgt, de_id, alpha, 1e-4, 1.0, 10000.0, paired)
    got = [r["loss_F"].item(), r["loss_gp"].item(), r["loss_T"].item(), r["loss_mse"].item()]
    want = [o["loss_F"], o["loss_gp"], o["loss_T"], o["loss_mse"]]
    print("losses", got, want)
    assert abs(got[0] - want[0]) < 5e-6
    assert abs(got[1] - want[1]) <= 2e-3 * abs(want[1])          # GP: after one sign-like RMSprop step on F
    # T(x) is computed before any update: the north-star tolerance applies to it directly, and the RMSE with it
    torch.testing.assert_close(r["out"].cpu(), o["out"], rtol=1e-3, atol=1e-4)
    assert abs(got[3] - want[3]) <= 2e-4 * abs(want[3])
    # Loss_T contains -mean f(T(x)) evaluated with the potential AFTER its two sign-like RMSprop steps (each weight
    # moves by +-10*lr whatever |g|; ~zero gradients flip sign under any change of summation order), so it carries
    # that step's noise: an absolute +-1..2 at P=128 whatever the size of the other terms (2e-4 relative at P=32,
    # tests/test_train_step.py)
    # (seven B200 runs per mode: paired 1704.9 -0.6 ... +0.8, unpaired 638.3 -1.5 ... +0.9 -- the noise is absolute, so
    # the bound is too: 5 is > 5 sigma of it and still 0.3 % / 0.8 % of the loss)
    assert abs(got[2] - want[2]) <= max(5e-3 * abs(want[2]), 5.0)
    # F-sub inside the iteration: L_F = mean f(fake) - mean f(real) ~ 5e-5 at initialisation -- its gradient is a ~1 %
    # residue of two cancelling terms, so the 1e-5-level difference between our T(x) and the oracle's shows up as ~1e-2
    _cmp_grads(Fp.ps, step.capture["F"], o["grads_F"], 3e-2, "F-sub (own T output)", tol_tensor=6e-2)
    # GP inside the iteration is evaluated at the potential's weights AFTER its first sign-like RMSprop step: every
    # weight whose F-sub gradient is ~0 may have moved +10*lr here and -10*lr in the oracle -> percent-level differences
    _cmp_grads(Fp.ps, step.capture["GP"], o["grads_GP"], 6e-2, "GP (own first F step)", tol_tensor=1.5e-1)
    # ... and the potential's backward itself, fed the ORACLE's fake batch at the initial weights: tight
    Ff = Fp.ps.flat.clone()
    Fp.ps.flat.copy_(Fflat0)
    Fp.ps.repack()
    Fp.ps.zero_grad()
    Fp.critic_step(tgt.cuda(), o["out"].cuda(), B)
    # (not tighter than the line above: what remains are single LeakyReLU-mask flips where a pre-activation is ~1e-9 --
    #  one flipped element of a 2.6e5-element delta moves its rel-L2 by 1.6e-3, and the real/fake cancellation
    #  amplifies that 3-6x; scripts/diag_fnet.py counts them.  Layers above the flip agree to 5e-5.)
    _cmp_grads(Fp.ps, Fp.ps.grad, o["grads_F"], 3e-2, "F-sub (oracle's T output)", tol_tensor=6e-2)
    # ... and the gradient penalty at IDENTICAL (initial) weights against double-backward autograd on the oracle: tight
    from rcot_b200 import ops
    Fl = {k: v.detach().clone().requires_grad_(True) for k, v in F0.items()}
    a4 = alpha.view(-1, 1, 1, 1)
    xt = (a4 * tgt + (1 - a4) * o["out"]).detach().requires_grad_(True)
    with torch.enable_grad():
        f = R.fnet_forward(Fl, xt)
        gx = torch.autograd.grad(f, xt, torch.ones_like(f), create_graph=True)[0]
        gp = 10 * ((gx.flatten(1).norm(dim=1) - 1) ** 2).mean()
        keys = list(Fl)
        gref = dict(zip(keys, torch.autograd.grad(gp, [Fl[k] for k in keys], allow_unused=True)))
    Fp.ps.zero_grad()
    lgp = Fp.penalty_step(ops.axpby(tgt.cuda(), o["out"].cuda(), a_vec=alpha.cuda()), B)
    assert abs(lgp.item() - gp.item()) <= 1e-4 * abs(gp.item())
    _cmp_grads(Fp.ps, Fp.ps.grad, {k: (None if k == "fc2.bias" else (torch.zeros_like(Fl[k]) if v is None else v))
                                   for k, v in gref.items()}, 3e-3, "GP (initial weights)", tol_tensor=1e-2)
    Fp.ps.flat.copy_(Ff)
    Fp.ps.repack()
    # ---- T-sub.  Inside the iteration dL/dout contains dF/dout of the potential AFTER its two sign-like steps (our
    # weights and the oracle's differ by +-10*lr on ~zero-gradient entries), so the transport gradient is checked in two
    # exact halves instead: (a) our cost kernels + F input gradient with the ORACLE's updated potential -> dL/dout;
    # (b) our T backward (initial weights) driven by the oracle's dL/dout -> every weight gradient.
    out_o = o["out"].clone().requires_grad_(True)
    with torch.enable_grad():
        lossT, _ = R.transport_loss(out_o, deg, tgt, R.fnet_forward(F_sd, out_o), de_id, 1.0, 10000.0, paired)
    lossT.backward()
    dref = out_o.grad
    for k, off in Fp.ps.offsets.items():
        Fp.ps.flat[off:off + F_sd[k].numel()].copy_(F_sd[k].flatten())
    Fp.ps.repack()
    out_g = o["out"].cuda()
    fv, dF = Fp.input_grad(out_g, -1.0 / B)
    acc = torch.zeros(4, device="cuda")
    gfou = torch.empty_like(out_g)
    tg = tgt.cuda() if paired else None
    ops.cost_stage1(out_g, deg.cuda(), tg, de_id.cuda(), gfou, acc)
    dout = torch.empty_like(out_g)
    ops.cost_stage2(out_g, deg.cuda(), tg, gfou, dF, acc, dout, 1.0, 10000.0, float(B * 3 * P * P))
    away = ((o["out"] - tgt).abs() > 1e-4) if paired else torch.ones_like(tgt, dtype=torch.bool)
    # (the Fourier |F| branch divides by |F|: bins with |F| ~ 0 are the only loose elements)
    err = (dout.cpu()[away] - dref[away]).norm() / dref[away].norm()
    print(f"dL/dout vs oracle (same potential): rel-L2 {err.item():.2e}")
    assert err < 2e-3, err
    Tp.ps.flat.copy_(Tflat0)
    Tp.ps.repack()
    Tp.ps.zero_grad()
    tape = Tape(save_hidden=True)
    out = Tp.forward(deg.cuda(), tape)
    tape.backward(out, dref.cuda())
    _cmp_grads(Tp.ps, Tp.ps.grad, o["grads_T"], 3e-3, f"T-sub ({'paired' if paired else 'unpaired'}, oracle dL/dout)")


def test_batch32_equals_sum_of_shards(cuda_lib):
    """Flat T / F gradient buffers at batch 32 (>148-tile persistent loops, hidden tensors saved, N slices) equal the
    sum over sixteen batch-2 shards run through the same kernels (fp32 summation order is the only difference)."""
    from rcot_b200 import ops
    from rcot_b200.engine import Tape
    P, B, S = 128, 32, 2
    Tp, Fp, _, _ = _nets(P)
    deg, tgt = _batch(5, B, P)
    deg, tgt = deg.cuda(), tgt.cuda()
    dout = torch.randn(B, 3, P, P, generator=torch.Generator().manual_seed(1)).cuda() * 1e-2
    alpha = torch.rand(B, generator=torch.Generator().manual_seed(2)).cuda()

    def t_grads(sl):
        Tp.ps.zero_grad()
        tape = Tape(save_hidden=True)
        out = Tp.forward(deg[sl].contiguous(), tape)
        tape.backward(out, dout[sl].clone())
        return Tp.ps.grad.clone(), out

    def f_grads(sl, out):
        Fp.ps.zero_grad()
        Fp.critic_step(tgt[sl].contiguous(), out, B)
        gc = Fp.ps.grad.clone()
        Fp.ps.zero_grad()
        interp = ops.axpby(tgt[sl].contiguous(), out, a_vec=alpha[sl].contiguous())
        Fp.penalty_step(interp, B)
        return gc, Fp.ps.grad.clone()

    full = slice(0, B)
    gT, out = t_grads(full)
    gF, gGP = f_grads(full, out)
    sT, sF, sGP = torch.zeros_like(gT), torch.zeros_like(gF), torch.zeros_like(gGP)
    for i in range(0, B, S):
        sl = slice(i, i + S)
        g, o = t_grads(sl)
        sT += g
        torch.testing.assert_close(o, out[sl], rtol=1e-4, atol=2e-5)     # MDTA's Gram atomics: summation order only
        a, b = f_grads(sl, o)
        sF += a
        sGP += b
    errs = {}
    for name, a, b in (("T", gT, sT), ("F", gF, sF), ("GP", gGP, sGP)):
        errs[name] = ((a - b).double().norm() / b.double().norm()).item()
        print(f"batch-32 vs 16 x batch-2, {name}: rel-L2 {errs[name]:.3e}")
    # Summation order is the only difference, but it is not a fixed one: the weight gradients and MDTA's Grams are
    # accumulated with atomics, so the figure moves from run to run (1.9e-5, 4.3e-5, 4.4e-5 in three runs of the same binary).
    # A wrong tile loop / N slice / saved tensor shows up as O(1e-2..1); the bounds leave 4x over the largest value seen.
    assert errs["T"] < 2e-4, errs
    # the critic gradient is a ~1 % residue of the cancelling real / fake contributions (L_F ~ 5e-5 at initialisation):
    # fp32 summation-order noise of either term shows up amplified by that ratio; the penalty gradient passes through
    # LeakyReLU sign masks, where ONE flipped element of a delta tensor moves that tensor by 1.6e-3 (scripts/diag_fnet.py)
    # (seen over the round's runs: F 5.5e-3, 5.9e-3, 1.0e-2; GP 6.6e-4, 9.4e-4)
    assert errs["F"] < 5e-2 and errs["GP"] < 5e-3, errs
