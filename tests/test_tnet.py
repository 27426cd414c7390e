"""GPU parity of the whole transport map and of the potential against (a) golden vectors from the
unmodified reference and (b) the CPU oracle on the same seeded inputs (P=32, B=2)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "rcot_golden.pt"), weights_only=False)


@pytest.fixture(scope="module")
def setup(gold, cuda_lib):
    import Net_Restormer as N
    from oracle.make_golden import synth_batch
    torch.manual_seed(0)
    T = N.T_net(decoder=True)
    F = N.F_net(patch_size=gold["P"])
    T_sd = {k: v.detach().clone() for k, v in T.state_dict().items()}
    F_sd = {k: v.detach().clone() for k, v in F.state_dict().items()}
    deg, tgt = synth_batch(1, gold["B"], gold["P"])
    return T.cuda(), F.cuda(), T_sd, F_sd, deg, tgt


def _stats(t):
    t = t.double()
    return torch.tensor([t.sum().item(), t.abs().sum().item(), t.norm().item()], dtype=torch.float64)


def _rel(name, got, ref, tol):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    err = (got - ref).norm().item() / max(ref.norm().item(), 1e-30)
    print(f"{name:48s} rel_l2_err={err:.3e}")
    assert err < tol, (name, err)


def test_tnet_forward_matches_reference_golden(gold, setup):
    T, F, *_ , deg, tgt = setup
    with torch.no_grad():
        out = T(deg.cuda())
    # north_star tolerance: rtol 1e-3 / atol 1e-4 in fp32
    torch.testing.assert_close(out.cpu(), gold["T_out"], rtol=1e-3, atol=1e-4)
    with torch.no_grad():
        torch.testing.assert_close(F(tgt.cuda()).cpu(), gold["F_tgt"], rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(F(gold["T_out"].cuda()).cpu(), gold["F_out"], rtol=1e-3, atol=1e-6)


def test_tnet_backward_matches_oracle_autograd(gold, setup):
    """Every one of the 796 gradient-carrying tensors, for a random upstream gradient."""
    from oracle import restormer_ref as R
    T, F, T_sd, F_sd, deg, tgt = setup
    g = torch.Generator().manual_seed(5)
    dout = torch.randn(deg.shape, generator=g)
    Tl = {k: v.clone().requires_grad_(True) for k, v in T_sd.items()}
    o = R.tnet_forward(Tl, deg)
    ref = dict(zip(Tl, torch.autograd.grad(o, list(Tl.values()), dout, allow_unused=True)))
    T.zero_grad()
    out = T(deg.cuda())
    out.backward(dout.cuda())
    n_none = 0
    for k, p in T.named_parameters():
        if ref[k] is None:
            assert p.grad is None, k
            n_none += 1
            continue
        _rel(k, p.grad, ref[k], 2e-3)
    assert n_none == 20      # 20 parameter tensors never receive a gradient


def test_fnet_objectives_match_reference_golden(gold, setup):
    from rcot_b200 import ops
    T, F, T_sd, F_sd, deg, tgt = setup
    B = gold["B"]
    with torch.no_grad():
        F(tgt.cuda())                      # builds the program
    prog = F._program
    out = gold["T_out"].cuda()
    prog.ps.zero_grad()
    loss = prog.critic_step(tgt.cuda(), out)
    torch.testing.assert_close(loss.cpu()[0], gold["loss_F"], rtol=1e-2, atol=1e-6)
    for k, w in gold["grads_F"].items():
        s = _stats(prog.gview(k))
        assert abs(s[2] - w[2]) <= 2e-3 * w[2] + 1e-10, (k, s, w)
        assert abs(s[0] - w[0]) <= 2e-3 * w[1] + 1e-10, (k, s, w)
    prog.ps.zero_grad()
    interp = ops.axpby(tgt.cuda(), out, a_vec=gold["alpha"].cuda())
    loss_gp = prog.penalty_step(interp)
    torch.testing.assert_close(loss_gp.cpu()[0], gold["loss_gp"], rtol=1e-3, atol=0)
    for k, w in gold["grads_GP"].items():
        s = _stats(prog.gview(k))
        if w is None or w[1] == 0:
            assert s[1] == 0, k            # zero / absent bias gradients stay exactly zero
            continue
        assert abs(s[2] - w[2]) <= 3e-3 * w[2] + 1e-10, (k, s, w)
        assert abs(s[0] - w[0]) <= 3e-3 * w[1] + 1e-10, (k, s, w)


def test_fnet_dropin_autograd(gold, setup):
    """loss.backward() through the drop-in F_net module (the reference trainer's usage)."""
    from oracle import restormer_ref as R
    T, F, T_sd, F_sd, deg, tgt = setup
    x = tgt.clone().requires_grad_(True)
    Fl = {k: v.clone().requires_grad_(True) for k, v in F_sd.items()}
    f = R.fnet_forward(Fl, x)
    w = torch.tensor([0.3, -1.1])
    (f * w).sum().backward()
    xd = tgt.cuda().requires_grad_(True)
    F.zero_grad()
    (F(xd) * w.cuda()).sum().backward()
    _rel("dx", xd.grad, x.grad, 2e-3)
    for k, p in F.named_parameters():
        _rel(k, p.grad, Fl[k].grad, 2e-3)


@pytest.mark.parametrize("H,W", [(40, 72), (128, 128), (24, 32)])
def test_tnet_whole_image_inference_matches_oracle(setup, H, W):
    """evaluate() / the testers run T_net on whole images of arbitrary size (multiples of 8; reference
    trainer.py:179-227, tester.py:62-114): feature maps such as 10x18 and 5x9 take the generic depthwise /
    register-prefetch GEMM paths, 128x128 the TMA-staged one.  Same weights, CPU oracle as the checker."""
    from oracle import restormer_ref as R
    T, _, T_sd, *_ = setup
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = torch.rand(1, 3, H, W, generator=g)
    with torch.no_grad():
        out = T(x.cuda())
        ref = R.tnet_forward(T_sd, x)
    assert out.shape == ref.shape == (1, 3, H, W)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-3, atol=1e-4)
