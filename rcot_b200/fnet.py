"""The OT potential F_net (reference Net_Restormer.py:436-522) and the three ways the training
step uses it (reference trainer.py:262-346), all with hand-derived backward passes:

* ``critic_step``   F-sub:  L = -mean f(real) + mean f(fake)  -> weight gradients
* ``penalty_step``  gradient penalty 10*mean_b(|grad_x f(x~)|-1)^2 WITHOUT double-backward autograd:
                    F_net is piecewise linear, so forward storing the LeakyReLU
                    signs, data-backward chain delta_l down to g = grad_x f, tangent-forward t_l of
                    u0 = dL/dg without biases, then dL/dW_l = wgrad(input=t_{l-1}, grad_out=delta_l).
                    Bias gradients are identically zero (fc2.bias gets none at all).
* ``input_grad``    T-sub:  d(-mean f(x))/dx for frozen weights.

Convolutions run on the tcgen05 pixel-as-M / pixel-as-K GEMMs; the three Linear layers on the
fp32 kernels of csrc/linear.cu.
"""
from __future__ import annotations

import torch

from . import ops
from .engine import ParamSet

CONVS = (  # (index in features, cin, cout, k, stride, pad, has_bias)
    (0, 3, 64, 5, 1, 2, True), (2, 64, 64, 4, 2, 1, True), (4, 64, 128, 3, 1, 1, True),
    (6, 128, 128, 4, 2, 1, True), (8, 128, 256, 3, 1, 1, True), (10, 256, 256, 4, 2, 1, True),
    (12, 256, 512, 3, 1, 1, False), (14, 512, 512, 4, 2, 1, False), (16, 512, 512, 3, 1, 1, False),
    (18, 512, 512, 4, 2, 1, False),
)
SLOPE = 0.2


class FnetProgram:
    def __init__(self, named_params, device, patch_size):
        if patch_size % 32:
            raise ValueError("F_net needs patch_size % 32 == 0")
        self.P = patch_size
        self.ps = ParamSet(named_params, device)
        self.grad_names = set(named_params)
        self.kf, self.kd = {}, {}
        for idx, *_ in CONVS:
            n = f"features.{idx}.weight"
            self.kf[idx] = ops.conv_pack_kind(self.ps.p[n], False)
            self.kd[idx] = ops.conv_pack_kind(self.ps.p[n], True)
            self.ps.add_pack(n, self.kf[idx])
            self.ps.add_pack(n, self.kd[idx])
        self.ps.finalize()
        # everything except fc2.bias (last tensor): the range the GP optimizer step covers
        self.n_without_fc2_bias = self.ps.offsets["fc2.bias"]

    def pview(self, name):
        return self.ps.p[name]

    def gview(self, name):
        return self.ps.g[name]

    # ------------------------------------------------------------------ forward
    def forward(self, x, bias=True, masks=None, t_in=None):
        """Returns (f [B], acts). acts = [x, y_1..y_10, h1, a2]: post-activation tensors.
        With ``masks`` (acts of a previous forward) runs the bias-free tangent pass instead:
        t_l = D_l (W_l * t_{l-1})."""
        ps = self.ps
        acts = [x]
        t = x
        for li, (idx, cin, cout, k, s, p, has_b) in enumerate(CONVS):
            w = f"features.{idx}."
            from . import engine
            if (li == 0 and engine.DIRECT_CONV3 and cin == 3 and s == 1 and k in (3, 5) and p == k // 2 and ops.TERMS == 3
                    and t.dtype == torch.float32):
                # features.0 (3 -> 64, 5x5): direct FP32 kernel instead of an implicit GEMM with K = 75
                if masks is None:
                    t = ops.conv_from3(t, ps.p[w + "weight"], bias=ps.p[w + "bias"] if has_b else None, act=True, slope=SLOPE)
                else:
                    t = ops.conv_from3(t, ps.p[w + "weight"], mask_y=masks[li + 1], slope=SLOPE)
            elif masks is None:
                t = ops.pm_gemm(t, ps.pack(w + "weight", self.kf[idx]), cout, ks=k, stride=s, pad=p,
                                bias=ps.p[w + "bias"] if has_b else None, act=True, slope=SLOPE,
                                tap_major=self.kf[idx].endswith("_tap"))
            else:
                t = ops.pm_gemm(t, ps.pack(w + "weight", self.kf[idx]), cout, ks=k, stride=s, pad=p,
                                mask_y=masks[li + 1], slope=SLOPE, tap_major=self.kf[idx].endswith("_tap"))
            acts.append(t)
        flat = t.view(t.shape[0], -1)
        if masks is None:
            h1 = ops.linear_fwd(flat, ps.p["fc.weight"], ps.p["fc.bias"])
            a2 = ops.linear_fwd(h1, ps.p["fc1.weight"], ps.p["fc1.bias"], act=True)
            f = ops.linear_fwd(a2, ps.p["fc2.weight"], ps.p["fc2.bias"]).view(-1)
        else:
            h1 = ops.linear_fwd(flat, ps.p["fc.weight"])
            a2 = ops.linear_fwd(h1, ps.p["fc1.weight"], mask=masks[12])
            f = None
        acts += [h1, a2]
        return f, acts

    # ------------------------------------------------------------------ backward chain
    def backward(self, acts, df, wgrad=True, need_dx=False, keep_deltas=False):
        """df: [B] = dL/df.  Accumulates weight/bias gradients (wgrad=True) into ps.grad and returns
        (dx or None, deltas or None); deltas = pre-activation gradients [conv1..conv10, fc, fc1, fc2]."""
        ps = self.ps
        x, h1, a2 = acts[0], acts[11], acts[12]
        y10 = acts[10]
        Bn = x.shape[0]
        d_f = df.view(Bn, 1).contiguous()
        if wgrad:
            ops.linear_wgrad(d_f, a2, ps.g["fc2.weight"], ps.g["fc2.bias"])
        d_h2 = ops.linear_dgrad(d_f, ps.p["fc2.weight"], mask=a2)
        if wgrad:
            ops.linear_wgrad(d_h2, h1, ps.g["fc1.weight"], ps.g["fc1.bias"])
        d_h1 = ops.linear_dgrad(d_h2, ps.p["fc1.weight"])
        flat = y10.view(Bn, -1)
        if wgrad:
            ops.linear_wgrad(d_h1, flat, ps.g["fc.weight"], ps.g["fc.bias"])
        delta = ops.linear_dgrad(d_h1, ps.p["fc.weight"], mask=flat).view(y10.shape)
        deltas = [None] * 10
        dx = None
        for li in range(9, -1, -1):
            idx, cin, cout, k, s, p, has_b = CONVS[li]
            w = f"features.{idx}."
            src = acts[li]
            deltas[li] = delta
            if wgrad:
                from . import engine
                if (li == 0 and engine.DIRECT_WGRAD3 and cin == 3 and s == 1 and k in (3, 5) and p == k // 2 and ops.TERMS == 3
                        and src.shape[3] <= 256):
                    ops.conv3_wgrad(delta, src, ps.g[w + "weight"], from3=True)     # features.0: direct FP32 kernel
                else:
                    ops.pk_gemm(delta, src, ps.g[w + "weight"].view(cout, -1), ldo=cin * k * k, ks=k, stride=s, pad=p)
                if has_b:
                    ops.channel_sum(delta, ps.g[w + "bias"])
            if li > 0:
                delta = ops.pm_gemm(delta, ps.pack(w + "weight", self.kd[idx]), cin, ks=k, stride=s, pad=p, mode=1,
                                    out_hw=(src.shape[2], src.shape[3]), mask_y=src, slope=SLOPE,
                                    tap_major=self.kd[idx].endswith("_tap"))
            elif need_dx:
                from . import engine
                if engine.DIRECT_CONV3 and cin == 3 and s == 1 and k in (3, 5) and p == k // 2 and ops.TERMS == 3:
                    dx = ops.conv_to3(delta, ps.p[w + "weight"], dgrad=True)   # 64 -> 3, 5x5: direct FP32 kernel
                else:
                    dx = ops.pm_gemm(delta, ps.pack(w + "weight", self.kd[idx]), cin, ks=k, stride=s, pad=p, mode=1,
                                     out_hw=(src.shape[2], src.shape[3]), tap_major=self.kd[idx].endswith("_tap"))
        return dx, (deltas + [d_h1, d_h2, d_f] if keep_deltas else None)

    # ------------------------------------------------------------------ the three uses
    def critic_step(self, real, fake, B_global=None):
        """Accumulates d/dW of -mean f(real) + mean f(fake); returns loss as a 1-element tensor."""
        B = real.shape[0]
        Bg = B if B_global is None else B_global
        x = torch.cat([real, fake], 0)
        f, acts = self.forward(x)
        df = torch.cat([torch.full((B,), -1.0 / Bg, device=x.device), torch.full((B,), 1.0 / Bg, device=x.device)])
        self.backward(acts, df, wgrad=True)
        loss = torch.zeros(1, device=x.device)
        ops.signed_sum(f, loss, B, 1.0 / Bg)
        return loss

    def penalty_step(self, interp, B_global=None):
        """Accumulates d/dW of 10*mean_b(|grad f(interp_b)| - 1)^2; returns the loss (1-element tensor)."""
        ps = self.ps
        B = interp.shape[0]
        Bg = B if B_global is None else B_global
        f, acts = self.forward(interp)
        g, deltas = self.backward(acts, torch.ones(B, device=interp.device), wgrad=False, need_dx=True,
                                  keep_deltas=True)
        stat = torch.zeros(2 * B + 1, device=interp.device)
        sumsq, coef, loss = stat[:B], stat[B:2 * B], stat[2 * B:]
        ops.sample_sumsq(g, sumsq)
        ops.gp_coef(sumsq, coef, loss, Bg)
        u0 = ops.axpby(g, None, a_vec=coef)
        _, tang = self.forward(u0, masks=acts)
        from . import engine
        for li, (idx, cin, cout, k, s, p, has_b) in enumerate(CONVS):
            if (li == 0 and engine.DIRECT_WGRAD3 and cin == 3 and s == 1 and k in (3, 5) and p == k // 2 and ops.TERMS == 3
                    and tang[li].shape[3] <= 256):
                ops.conv3_wgrad(deltas[li], tang[li], ps.g[f"features.{idx}.weight"], from3=True)
                continue
            ops.pk_gemm(deltas[li], tang[li], ps.g[f"features.{idx}.weight"].view(cout, -1), ldo=cin * k * k, ks=k,
                        stride=s, pad=p)
        d_h1, d_h2, d_f = deltas[10], deltas[11], deltas[12]
        ops.linear_wgrad(d_h1, tang[10].view(B, -1), ps.g["fc.weight"])
        ops.linear_wgrad(d_h2, tang[11], ps.g["fc1.weight"])
        ops.linear_wgrad(d_f, tang[12], ps.g["fc2.weight"])
        return loss

    def input_grad(self, x, scale):
        """Returns (f [B], d(scale * sum_b f_b)/dx) for frozen weights."""
        f, acts = self.forward(x)
        dx, _ = self.backward(acts, torch.full((x.shape[0],), scale, device=x.device), wgrad=False, need_dx=True)
        return f, dx

    # ------------------------------------------------------------------ drop-in autograd bridge
    def forward_tape(self, x, tape):
        f, acts = self.forward(x)
        if tape.enabled:
            def bwd(df):
                dx, _ = self.backward(acts, df, wgrad=True, need_dx=True)
                tape.add_grad(x, dx)
            tape.record(f, bwd)
        return f
