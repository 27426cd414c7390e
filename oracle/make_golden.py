"""ORACLE tooling (this container only): run the UNMODIFIED reference from /root/reference on CPU
and commit small golden vectors to tests/golden/rcot_golden.pt.

    python -m oracle.make_golden

The reference has no tests or known-answer vectors of its own, so parity is
pinned by executing it: outputs of its T_net / F_net classes, gradients of the three objectives of
one trainer.py iteration obtained with ITS autograd graph (per-tensor sum / abs-sum / L2 norm), and
the losses its own ``trainer.train()`` prints.  Weights are not stored: both the reference and the
drop-in modules reproduce them from ``torch.manual_seed`` (checked by checksum).
"""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "rcot_golden.pt")


def synth_batch(seed, B, P):
    """Synthetic paired batch: target U[0,1), degraded = target + sigma 25/255 noise."""
    g = torch.Generator().manual_seed(seed)
    tgt = torch.rand(B, 3, P, P, generator=g)
    deg = tgt + 25.0 / 255.0 * torch.randn(B, 3, P, P, generator=g)
    return deg, tgt


def stats(t):
    t = t.double()
    return torch.tensor([t.sum().item(), t.abs().sum().item(), t.norm().item()], dtype=torch.float64)


def grad_stats(module):
    return {k: (None if p.grad is None else stats(p.grad)) for k, p in module.named_parameters()}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    net = ref_shim.import_net()
    G = {"torch": torch.__version__}
    P, B = 32, 2
    torch.manual_seed(0)
    T = net.T_net(decoder=True)
    F = net.F_net(patch_size=P)
    G["param_sum_T"] = stats(torch.cat([p.flatten() for p in T.parameters()]))
    G["param_sum_F"] = stats(torch.cat([p.flatten() for p in F.parameters()]))
    deg, tgt = synth_batch(1, B, P)
    de_id = torch.tensor([1, 4])
    G["P"], G["B"], G["de_id"] = P, B, de_id

    # ---- forward values
    with torch.no_grad():
        out = T(deg)
        G["T_out"] = out.clone()
        G["F_tgt"] = F(tgt).clone()
        G["F_out"] = F(out).clone()
    # ---- F-sub gradients (trainer.py:268-276)
    F.zero_grad()
    loss_F = -F(tgt).squeeze().mean() + F(out).squeeze().mean()
    loss_F.backward()
    G["loss_F"] = loss_F.detach()
    G["grads_F"] = grad_stats(F)
    # ---- gradient penalty gradients (trainer.py:283-307), same weights
    F.zero_grad()
    alpha = torch.rand(B, 1, 1, 1, generator=torch.Generator().manual_seed(7))
    G["alpha"] = alpha.flatten().clone()
    inter = (alpha.expand_as(tgt) * tgt + (1 - alpha.expand_as(tgt)) * out).requires_grad_(True)
    o = F(inter).squeeze()
    grad = torch.autograd.grad(outputs=o, inputs=inter, grad_outputs=torch.ones(o.size()), retain_graph=True,
                               create_graph=True, only_inputs=True)[0]
    gp = 10 * torch.mean((torch.sqrt(torch.sum(grad.view(B, -1) ** 2, dim=1)) - 1) ** 2)
    gp.backward()
    G["loss_gp"] = gp.detach()
    G["grads_GP"] = grad_stats(F)
    G["gp_input_grad"] = grad.detach().clone()
    # ---- T-sub gradients (trainer.py:318-345), paired branch, both Fourier branches in the batch
    for p in F.parameters():
        p.requires_grad_(False)
    T.zero_grad()
    out = T(deg)
    out_disc = F(out).squeeze()
    res = deg - out
    mse_loss = (torch.mean(res ** 2)) ** 0.5
    res_fre = torch.fft.fft2(res)
    fourier = 0
    for i in range(B):
        sl = res_fre[i, :]
        if de_id[i] < 3:
            fourier += torch.mean(abs(sl) ** 2) ** 1 / 2
        else:
            fourier += torch.mean(abs(sl))
    sigma, Sigma = 1.0, 10000.0
    T_loss = -out_disc.mean() + sigma * (mse_loss + fourier) + Sigma * torch.mean(abs(out - tgt))
    T_loss.backward()
    G["loss_T"], G["loss_mse"], G["fourier"] = T_loss.detach(), mse_loss.detach(), fourier.detach()
    G["grads_T"] = grad_stats(T)

    # ---- the reference's own train() for one iteration (printed losses + post-step checksums)
    with tempfile.TemporaryDirectory() as wd:
        cwd = os.getcwd()
        os.chdir(wd)
        try:
            tr, net2 = ref_shim.import_trainer(wd, ["--batchSize", str(B), "--patch_size", str(P), "--pairnum", "10000000",
                                                    "--type", "Denoising"])
            torch.manual_seed(0)
            T2 = net2.T_net(decoder=True)
            F2 = net2.F_net(patch_size=P)
            T_opt = torch.optim.RMSprop(T2.parameters(), lr=tr.opt.lr / 2)
            F_opt = torch.optim.RMSprop(F2.parameters(), lr=tr.opt.lr)
            loader = [([["a", "b"], de_id], deg, tgt)]
            torch.manual_seed(123)      # fixes the alpha draw of the gradient penalty
            G["train_alpha"] = torch.rand(B, 1, 1, 1).flatten().clone()
            torch.manual_seed(123)
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                tr.train(loader, T_opt, F_opt, T2, F2, 1)
            m = re.search(r"Loss_F: ([-\de.+]+), Loss_T: ([-\de.+]+), Loss_mse: ([-\de.+]+)", buf.getvalue())
            G["train_losses"] = torch.tensor([float(m.group(i)) for i in (1, 2, 3)], dtype=torch.float64)
            G["train_param_sum_T"] = stats(torch.cat([p.detach().flatten() for p in T2.parameters()]))
            G["train_param_sum_F"] = stats(torch.cat([p.detach().flatten() for p in F2.parameters()]))
            # a second pass of the updated nets pins the post-step weights through their outputs
            with torch.no_grad():
                G["train_T_out_after"] = T2(deg).clone()
                G["train_F_tgt_after"] = F2(tgt).clone()
        finally:
            os.chdir(cwd)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(G, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    for k in ("loss_F", "loss_gp", "loss_T", "loss_mse", "fourier", "train_losses"):
        print(k, G[k])


if __name__ == "__main__":
    main()
