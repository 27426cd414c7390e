"""ORACLE (test infrastructure only).  CPU restatement of one iteration of the reference's
adversarial loop -- /root/reference/trainer.py:247-346 -- on top of oracle.restormer_ref:

  F-sub  : L_F = -mean f(target) + mean f(T(x).detach()); RMSprop(lr)            (:262-280)
  GP     : 10 * mean_b (||grad_xt f(xt)||_2 - 1)^2, xt = a*target + (1-a)*fake;   (:283-308)
           second RMSprop(lr) step on F
  T-sub  : L_T = -mean f(T(x)) + sigma*(rmse + fourier) [+ Sigma*L1 if paired];   (:311-346)
           RMSprop(lr/2) on T

RMSprop follows torch.optim.RMSprop defaults (alpha .99, eps 1e-8, no momentum, not centered):
parameters whose gradient is None are skipped, zero gradients still decay square_avg.
"""
from __future__ import annotations

import torch

from . import restormer_ref as R


def rmsprop_step(params, grads, state, lr, alpha=0.99, eps=1e-8):
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        sq = state.setdefault(k, torch.zeros_like(p))
        sq.mul_(alpha).addcmul_(g, g, value=1 - alpha)
        p.addcdiv_(g, sq.sqrt().add_(eps), value=-lr)


def _leaves(sd):
    return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}


def _grads(loss, leaves):
    keys = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
    return dict(zip(keys, gs))


def train_iteration(T_sd, F_sd, T_state, F_state, degraded, target, de_id, alpha_gp, lr,
                    sigma=1.0, Sigma=10000.0, paired=True):
    """Runs one iteration in place on the tensors of T_sd / F_sd.  Returns the three printed losses."""
    with torch.no_grad():
        fake = R.tnet_forward(T_sd, degraded)
    # ---- F-sub
    Fl = _leaves(F_sd)
    loss_F = -R.fnet_forward(Fl, target).mean() + R.fnet_forward(Fl, fake).mean()
    gF = _grads(loss_F, Fl)
    with torch.no_grad():
        rmsprop_step(F_sd, gF, F_state, lr)
    # ---- gradient penalty (second F step, on the updated weights)
    Fl = _leaves(F_sd)
    a = alpha_gp.view(-1, 1, 1, 1).to(target)
    xt = (a * target + (1 - a) * fake).detach().requires_grad_(True)
    f = R.fnet_forward(Fl, xt)
    g = torch.autograd.grad(f, xt, torch.ones_like(f), create_graph=True)[0]
    gp = 10 * ((g.flatten(1).norm(dim=1) - 1) ** 2).mean()
    gGP = _grads(gp, Fl)
    with torch.no_grad():
        rmsprop_step(F_sd, gGP, F_state, lr)
    # ---- T-sub
    Tl = _leaves(T_sd)
    out = R.tnet_forward(Tl, degraded)
    f_out = R.fnet_forward(F_sd, out)
    loss_T, rmse = R.transport_loss(out, degraded, target, f_out, de_id, sigma, Sigma, paired)
    gT = _grads(loss_T, Tl)
    with torch.no_grad():
        rmsprop_step(T_sd, gT, T_state, lr / 2)
    return {"loss_F": float(loss_F), "loss_gp": float(gp), "loss_T": float(loss_T), "loss_mse": float(rmse),
            "grads_T": gT, "grads_F": gF, "grads_GP": gGP, "out": out.detach()}
