"""Drop-in boundary details the reference's callers rely on: a stand-alone LayerNorm module, the gradient
to T_net's INPUT when it requires grad, `train()` driven by the torch.optim objects the reference's main() builds
(trainer.py:121-126), and a loud error (not silent garbage) for double backward through the hand-derived backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_layernorm_standalone_forward_backward(cuda_lib):
    import Net_Restormer as N
    torch.manual_seed(0)
    ln = N.LayerNorm(48, 'WithBias').cuda()
    with torch.no_grad():
        ln.body.weight.uniform_(0.5, 1.5)
        ln.body.bias.uniform_(-0.5, 0.5)
    w, b = ln.body.weight.detach().double().cpu(), ln.body.bias.detach().double().cpu()
    x = torch.randn(2, 48, 12, 20, device="cuda", requires_grad=True)
    y = ln(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    xd = x.detach().double().cpu().requires_grad_(True)
    x3 = xd.permute(0, 2, 3, 1)                                   # to_3d / to_4d of the reference (:96-101,198-200)
    mu, var = x3.mean(-1, keepdim=True), x3.var(-1, keepdim=True, unbiased=False)
    wd, bd = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = (((x3 - mu) / torch.sqrt(var + 1e-5)) * wd + bd).permute(0, 3, 1, 2)
    yr.backward(dy.double().cpu())
    torch.testing.assert_close(y.detach().cpu().double(), yr.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad.cpu().double(), xd.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ln.body.weight.grad.cpu().double(), wd.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ln.body.bias.grad.cpu().double(), bd.grad, rtol=1e-4, atol=1e-4)


def test_tnet_input_gradient_matches_oracle(cuda_lib):
    import Net_Restormer as N
    from oracle import restormer_ref as R
    P = 32
    torch.manual_seed(0)
    T = N.T_net(decoder=True)
    T_sd = {k: v.detach().clone() for k, v in T.state_dict().items()}
    T = T.cuda()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, P, P, generator=g)
    dout = torch.randn(1, 3, P, P, generator=g)
    xo = x.clone().requires_grad_(True)
    R.tnet_forward(T_sd, xo).backward(dout)
    xg = x.cuda().requires_grad_(True)
    T(xg).backward(dout.cuda())
    assert xg.grad is not None
    err = (xg.grad.cpu() - xo.grad).norm() / xo.grad.norm()
    assert err < 2e-3, err


def test_double_backward_raises(cuda_lib):
    import Net_Restormer as N
    torch.manual_seed(0)
    F = N.F_net(patch_size=32).cuda()
    x = torch.rand(2, 3, 32, 32, device="cuda", requires_grad=True)
    out = F(x).squeeze()
    g = torch.autograd.grad(out, x, torch.ones_like(out), create_graph=True)[0]
    with pytest.raises(RuntimeError):
        (g.flatten(1).norm(dim=1) - 1).pow(2).mean().backward()


def test_train_accepts_reference_torch_optim(cuda_lib):
    """trainer.train() with torch.optim.RMSprop objects (what the reference's main() passes) takes the same steps as with
    the EngineOptimizer facade, and writes the scheduled learning rates into their param_groups like the reference."""
    import Net_Restormer as N
    import trainer
    P, B = 32, 2
    trainer.opt = trainer.parser.parse_args(["--patch_size", str(P), "--batchSize", str(B), "--no_dump", "--pairnum", "100"])
    g = torch.Generator().manual_seed(2)
    tgt = torch.rand(B, 3, P, P, generator=g)
    deg = tgt + 0.1 * torch.randn(B, 3, P, P, generator=g)
    loader = [([["a", "b"], torch.tensor([1, 3])], deg, tgt)] * 2
    res = []
    for use_torch in (True, False):
        torch.manual_seed(0)
        T, F = N.T_net(decoder=True).cuda(), N.F_net(patch_size=P).cuda()
        if use_torch:
            To = torch.optim.RMSprop(T.parameters(), lr=trainer.opt.lr / 2)
            Fo = torch.optim.RMSprop(F.parameters(), lr=trainer.opt.lr)
        else:
            To, Fo = trainer.EngineOptimizer("RMSprop", trainer.opt.lr / 2), trainer.EngineOptimizer("RMSprop", trainer.opt.lr)
        torch.manual_seed(9)                     # alpha draws
        trainer.train(loader, To, Fo, T, F, 21)  # epoch 21 -> lr * 0.1
        assert abs(Fo.param_groups[0]["lr"] - trainer.opt.lr * 0.1) < 1e-12
        assert abs(To.param_groups[0]["lr"] - trainer.opt.lr * 0.05) < 1e-12
        with torch.no_grad():
            res.append(T(deg.cuda()).cpu())
    torch.testing.assert_close(res[0], res[1], rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        trainer._optimizer_kind(torch.optim.RMSprop(F.parameters(), lr=1e-4, momentum=0.9))
    with pytest.raises(TypeError):
        trainer._optimizer_kind(torch.optim.SGD(F.parameters(), lr=1e-4))
