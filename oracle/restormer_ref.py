"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement, in plain functional PyTorch, of the reference's transport map and potential:

* LayerNorm over channels ........ /root/reference/Net_Restormer.py:173-200
* MDTA channel attention ......... /root/reference/Net_Restormer.py:19-50
* GDFN gated feed-forward ........ /root/reference/Net_Restormer.py:67-85
* TransformerBlock ............... /root/reference/Net_Restormer.py:201-214
* Down/Upsample, patch embed ..... /root/reference/Net_Restormer.py:86-122
* T_net two-pass forward ......... /root/reference/Net_Restormer.py:328-434 (decoder=True)
* F_net potential ................ /root/reference/Net_Restormer.py:436-522

Every function takes a ``state_dict``-style mapping with the reference's key names, so the same
weights drive the reference, this oracle and the CUDA path.  Works in fp32 or fp64 and is
differentiable with autograd (the oracle for the hand-derived backward kernels).

Parity pinning: validated against the imported reference (tests/test_oracle_vs_reference.py, runs
where /root/reference exists) and against the committed golden vectors in tests/golden/ that
oracle/make_golden.py generated from the unmodified reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

DIM = 48
NUM_BLOCKS = (4, 6, 6, 8)
HEADS = (1, 2, 4, 8)
FFN_FACTOR = 2.66


def layer_norm_c(x, weight, bias):
    """Per-pixel LayerNorm over the channel axis of NCHW (biased variance, eps 1e-5)."""
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def mdta(x, sd, pre, heads):
    b, c, h, w = x.shape
    qkv = F.conv2d(x, sd[pre + "qkv.weight"])
    qkv = F.conv2d(qkv, sd[pre + "qkv_dwconv.weight"], padding=1, groups=3 * c)
    q, k, v = qkv.view(b, 3, heads, c // heads, h * w).unbind(1)
    q = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    k = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    attn = torch.softmax((q @ k.transpose(-1, -2)) * sd[pre + "temperature"].view(1, heads, 1, 1), dim=-1)
    out = (attn @ v).reshape(b, c, h, w)
    return F.conv2d(out, sd[pre + "project_out.weight"])


def gdfn(x, sd, pre):
    u = F.conv2d(x, sd[pre + "project_in.weight"])
    u = F.conv2d(u, sd[pre + "dwconv.weight"], padding=1, groups=u.shape[1])
    a, g = u.chunk(2, dim=1)
    return F.conv2d(F.gelu(a) * g, sd[pre + "project_out.weight"])


def transformer_block(x, sd, pre, heads):
    x = x + mdta(layer_norm_c(x, sd[pre + "norm1.body.weight"], sd[pre + "norm1.body.bias"]), sd, pre + "attn.", heads)
    x = x + gdfn(layer_norm_c(x, sd[pre + "norm2.body.weight"], sd[pre + "norm2.body.bias"]), sd, pre + "ffn.")
    return x


def _stage(x, sd, name, n, heads):
    for i in range(n):
        x = transformer_block(x, sd, f"{name}.{i}.", heads)
    return x


def _down(x, sd, name):
    return F.pixel_unshuffle(F.conv2d(x, sd[name + ".body.0.weight"], padding=1), 2)


def _up(x, sd, name):
    return F.pixel_shuffle(F.conv2d(x, sd[name + ".body.0.weight"], padding=1), 2)


def _decode(latent, skips, sd, img):
    """Shared decoder tail used by both passes (reference lines 345-375 and 400-432)."""
    e1, e2, e3 = skips
    t = transformer_block(latent, sd, "noise_level3.", HEADS[2])
    t = F.conv2d(t, sd["reduce_noise_level3.weight"])
    t = _up(t, sd, "up4_3")
    t = F.conv2d(torch.cat([t, e3], 1), sd["reduce_chan_level3.weight"])
    t = _stage(t, sd, "decoder_level3", NUM_BLOCKS[2], HEADS[2])
    t = transformer_block(t, sd, "noise_level2.", HEADS[2])
    t = F.conv2d(t, sd["reduce_noise_level2.weight"])
    t = _up(t, sd, "up3_2")
    t = F.conv2d(torch.cat([t, e2], 1), sd["reduce_chan_level2.weight"])
    t = _stage(t, sd, "decoder_level2", NUM_BLOCKS[1], HEADS[1])
    t = transformer_block(t, sd, "noise_level1.", HEADS[2])
    t = F.conv2d(t, sd["reduce_noise_level1.weight"])
    t = _up(t, sd, "up2_1")
    t = torch.cat([t, e1], 1)
    t = _stage(t, sd, "decoder_level1", NUM_BLOCKS[0], HEADS[0])
    t = _stage(t, sd, "refinement", 4, HEADS[0])
    return F.conv2d(t, sd["output.weight"], padding=1) + img


def tnet_forward(sd, img, return_residual=False):
    """Two-pass transport map with the residual conditioner on the latent (decoder=True)."""
    e1 = _stage(F.conv2d(img, sd["patch_embed.proj.weight"], padding=1), sd, "encoder_level1", NUM_BLOCKS[0], HEADS[0])
    e2 = _stage(_down(e1, sd, "down1_2"), sd, "encoder_level2", NUM_BLOCKS[1], HEADS[1])
    e3 = _stage(_down(e2, sd, "down2_3"), sd, "encoder_level3", NUM_BLOCKS[2], HEADS[2])
    l4 = _down(e3, sd, "down3_4")
    latent = _stage(l4, sd, "latent", NUM_BLOCKS[3], HEADS[3])
    first = _decode(latent, (e1, e2, e3), sd, img)
    res = img - first
    # residual conditioner: shares patch_embed and down3_4 with the main path (reference :381,:393)
    r = _stage(F.conv2d(res, sd["patch_embed.proj.weight"], padding=1), sd, "resencoder_level1", NUM_BLOCKS[0], HEADS[0])
    r = _stage(_down(r, sd, "resdown1_2"), sd, "resencoder_level2", NUM_BLOCKS[1], HEADS[1])
    r = _stage(_down(r, sd, "resdown2_3"), sd, "resencoder_level3", NUM_BLOCKS[2], HEADS[2])
    r = _stage(_down(r, sd, "down3_4"), sd, "reslatent", NUM_BLOCKS[3], HEADS[3])
    # the reference recomputes latent(l4) here (:397); it is value-identical to `latent`
    out = _decode(latent + 0.8 * r, (e1, e2, e3), sd, img)
    return (out, res) if return_residual else out


FNET_CONVS = (  # (index in features, cin, cout, k, stride, pad, has_bias)
    (0, 3, 64, 5, 1, 2, True), (2, 64, 64, 4, 2, 1, True), (4, 64, 128, 3, 1, 1, True),
    (6, 128, 128, 4, 2, 1, True), (8, 128, 256, 3, 1, 1, True), (10, 256, 256, 4, 2, 1, True),
    (12, 256, 512, 3, 1, 1, False), (14, 512, 512, 4, 2, 1, False), (16, 512, 512, 3, 1, 1, False),
    (18, 512, 512, 4, 2, 1, False),
)


def fnet_forward(sd, x):
    """Potential f: 10 x (conv + LeakyReLU 0.2), flatten, fc, fc1, LeakyReLU, fc2 -> [B]."""
    for idx, _, _, _, s, p, has_b in FNET_CONVS:
        x = F.leaky_relu(F.conv2d(x, sd[f"features.{idx}.weight"], sd.get(f"features.{idx}.bias") if has_b else None,
                                  stride=s, padding=p), 0.2)
    x = x.flatten(1)
    x = F.linear(x, sd["fc.weight"], sd["fc.bias"])          # no activation between fc and fc1
    x = F.leaky_relu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]), 0.2)
    return F.linear(x, sd["fc2.weight"], sd["fc2.bias"]).view(-1)


def fourier_cost(res, de_id):
    """Sum over the batch of the per-sample Fourier penalty (trainer.py:323-332).

    ``mean(|F|**2)**1/2`` parses as ``mean(|F|^2) / 2`` in the reference; kept.
    """
    spec = torch.fft.fft2(res)
    total = res.new_zeros(())
    for i in range(res.shape[0]):
        mag = spec[i].abs()
        total = total + ((mag ** 2).mean() / 2 if int(de_id[i]) < 3 else mag.mean())
    return total


def transport_loss(out_restored, degraded, target, f_out, de_id, sigma, Sigma, paired):
    """T-sub objective (trainer.py:319-343). Returns (loss, rmse)."""
    res = degraded - out_restored
    rmse = torch.sqrt((res ** 2).mean())
    loss = -f_out.mean() + sigma * (rmse + fourier_cost(res, de_id))
    if paired:
        loss = loss + Sigma * (out_restored - target).abs().mean()
    return loss, rmse


def gelu_exact(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
