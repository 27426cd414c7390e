#!/usr/bin/env python
"""Diagnostic: per-tensor error of the potential's weight gradients (one-sided objective mean f(x), no cancellation)
and of its input gradient, GPU kernels vs the CPU oracle, at P = 32 / 64 / 128."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import Net_Restormer as N  # noqa: E402
from oracle import restormer_ref as R  # noqa: E402
from rcot_b200.fnet import FnetProgram  # noqa: E402

for P in (32, 64, 128):
    torch.manual_seed(0)
    F = N.F_net(patch_size=P)
    F_sd = {k: v.detach().clone() for k, v in F.state_dict().items()}
    Fp = FnetProgram({k: v.detach().cuda() for k, v in F.named_parameters()}, "cuda", P)
    g = torch.Generator().manual_seed(3)
    B = 2
    x = torch.rand(B, 3, P, P, generator=g)
    Fl = {k: v.clone().requires_grad_(True) for k, v in F_sd.items()}
    xo = x.clone().requires_grad_(True)
    f = R.fnet_forward(Fl, xo)
    keys = list(Fl)
    gr = torch.autograd.grad(f.mean(), [Fl[k] for k in keys] + [xo], allow_unused=True)
    ref = dict(zip(keys, gr[:-1]))
    # GPU: forward + backward with df = 1/B for every sample
    Fp.ps.zero_grad()
    fo, acts = Fp.forward(x.cuda())
    dx, _ = Fp.backward(acts, torch.full((B,), 1.0 / B, device="cuda"), wgrad=True, need_dx=True)
    # LeakyReLU mask agreement with the oracle, layer by layer (a single flipped element explains a 1e-3 rel-L2 jump)
    with torch.no_grad():
        t = x.clone()
        from oracle.restormer_ref import FNET_CONVS
        import torch.nn.functional as TF
        for li, (idx, _, _, _, s_, p_, has_b) in enumerate(FNET_CONVS):
            pre = TF.conv2d(t, F_sd[f"features.{idx}.weight"], F_sd.get(f"features.{idx}.bias") if has_b else None,
                            stride=s_, padding=p_)
            t = TF.leaky_relu(pre, 0.2)
            mism = ((acts[li + 1].cpu() > 0) != (pre > 0)).sum().item()
            if mism:
                bad = ((acts[li + 1].cpu() > 0) != (pre > 0))
                print(f"   layer {idx}: {mism} mask flips of {pre.numel()} (|pre| there: {pre[bad].abs().max().item():.1e})")
    print(f"P={P}: f err {(fo.cpu() - f.detach()).abs().max().item():.2e} (|f| {f.abs().max().item():.2e}); "
          f"dx rel {((dx.cpu() - gr[-1]).norm() / gr[-1].norm()).item():.2e}")
    for k in keys:
        if ref[k] is None:
            continue
        got = Fp.ps.g[k].cpu()
        print(f"   {k:22s} rel-L2 {((got - ref[k]).norm() / ref[k].norm()).item():.2e}  |g| {ref[k].norm().item():.2e}")
