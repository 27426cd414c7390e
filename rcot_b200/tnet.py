"""The two-pass Restormer transport map as a program over engine ops.

Follows reference Net_Restormer.py:328-434 (decoder=True): pass 1 -> residual -> residual encoder
-> latent += 0.8 * reslatent -> pass 2, with the modules shared between the passes invoked twice
(their gradients accumulate) and the quirks of the reference kept: the residual branch reuses
``patch_embed`` and ``down3_4`` (:381,:393), ``res_patch_embed``/``resdown3_4``/... never run.
The one deliberate deviation is value-neutral: ``self.latent(inp_enc_level4)`` is evaluated once
(the reference evaluates it a second time at :397 with identical inputs and weights).
"""
from __future__ import annotations

import torch

from . import engine, ops
from .engine import BlockSpec, ConvSpec, ParamSet, Tape

DIM = 48
NUM_BLOCKS = (4, 6, 6, 8)
HEADS = (1, 2, 4, 8)

# (module name, number of blocks, channels, heads) of every block stack that T_net.forward runs
STAGES = (
    ("encoder_level1", 4, 48, 1), ("encoder_level2", 6, 96, 2), ("encoder_level3", 6, 192, 4),
    ("latent", 8, 384, 8),
    ("resencoder_level1", 4, 48, 1), ("resencoder_level2", 6, 96, 2), ("resencoder_level3", 6, 192, 4),
    ("reslatent", 8, 384, 8),
    ("decoder_level3", 6, 192, 4), ("decoder_level2", 6, 96, 2), ("decoder_level1", 4, 96, 1),
    ("refinement", 4, 96, 1),
)
SINGLE_BLOCKS = (("noise_level3", 384, 4), ("noise_level2", 192, 4), ("noise_level1", 96, 4))
CONV3 = ("patch_embed.proj.weight", "down1_2.body.0.weight", "down2_3.body.0.weight", "down3_4.body.0.weight",
         "resdown1_2.body.0.weight", "resdown2_3.body.0.weight", "up4_3.body.0.weight", "up3_2.body.0.weight",
         "up2_1.body.0.weight", "output.weight")
CONV1 = ("reduce_noise_level3.weight", "reduce_noise_level2.weight", "reduce_noise_level1.weight",
         "reduce_chan_level3.weight", "reduce_chan_level2.weight")
# top-level modules that exist in the state_dict but never run (no gradient; reference :232-292)
UNUSED_PREFIXES = ("res_patch_embed.", "chnl_reduce1.", "chnl_reduce2.", "chnl_reduce3.", "reduce_noise_channel_1.",
                   "reduce_noise_channel_2.", "reduce_noise_channel_3.", "resdown3_4.", "resnoise_level3.",
                   "resreduce_noise_level3.")


def used_names(names):
    return {n for n in names if not n.startswith(UNUSED_PREFIXES)}


# Gradient buckets in the order the backward FINISHES them (modules shared by the two passes accumulate until their
# pass-1 use has been differentiated): 0 = residual conditioner, 1 = decoder side (both passes), 2 = latent,
# 3 = encoder + patch_embed + down3_4 (last).  TnetProgram.forward places tape markers at the matching points.
_BUCKET0 = ("resencoder_level", "reslatent.", "resdown1_2.", "resdown2_3.")
_BUCKET1 = ("noise_level", "reduce_noise_level", "up4_3.", "up3_2.", "up2_1.", "reduce_chan_level", "decoder_level",
            "refinement.", "output.")


def bucket_of(name):
    if name.startswith(_BUCKET0):
        return 0
    if name.startswith(_BUCKET1):
        return 1
    if name.startswith("latent."):
        return 2
    return 3


class TnetProgram:
    def __init__(self, named_params, device):
        self.ps = ParamSet(named_params, device, used=used_names(named_params), bucket_of=bucket_of)
        ps = self.ps
        self.stages = {name: [BlockSpec(ps, f"{name}.{i}.", C, h) for i in range(n)] for name, n, C, h in STAGES}
        self.single = {name: BlockSpec(ps, name + ".", C, h) for name, C, h in SINGLE_BLOCKS}
        self.conv = {}
        for name in CONV3:
            self.conv[name.split(".")[0]] = ConvSpec(ps, name, 3, 1)
        for name in CONV1:
            self.conv[name.split(".")[0]] = ConvSpec(ps, name, 1, 0)
        ps.finalize()

        self.grad_names = used_names(named_params)

    def pview(self, name):
        return self.ps.p[name]

    def gview(self, name):
        return self.ps.g[name]

    # ------------------------------------------------------------------ pieces
    def _stage(self, x, name, tape):
        for bs in self.stages[name]:
            x = engine.block_fwd(bs, x, tape)
        return x

    def _down(self, x, name, tape):
        return engine.shuffle_fwd(engine.conv_fwd(self.conv[name], x, tape), True, tape)

    def _up(self, x, name, tape, out=None):
        return engine.shuffle_fwd(engine.conv_fwd(self.conv[name], x, tape), False, tape, out=out)

    def _decode(self, latent, e1, e2, e3, img, tape, input_grad=False):
        cv = self.conv
        t = engine.block_fwd(self.single["noise_level3"], latent, tape)
        t = engine.conv_fwd(cv["reduce_noise_level3"], t, tape)
        t = self._up(t, "up4_3", tape)
        t = engine.conv_fwd(cv["reduce_chan_level3"], t, tape, x2=e3)
        t = self._stage(t, "decoder_level3", tape)
        t = engine.block_fwd(self.single["noise_level2"], t, tape)
        t = engine.conv_fwd(cv["reduce_noise_level2"], t, tape)
        t = self._up(t, "up3_2", tape)
        t = engine.conv_fwd(cv["reduce_chan_level2"], t, tape, x2=e2)
        t = self._stage(t, "decoder_level2", tape)
        t = engine.block_fwd(self.single["noise_level1"], t, tape)
        t = engine.conv_fwd(cv["reduce_noise_level1"], t, tape)
        t = engine.conv_fwd(cv["up2_1"], t, tape)
        t = engine.shuffle_cat_fwd(t, e1, tape)            # PixelShuffle + cat([., e1]) in one buffer
        t = self._stage(t, "decoder_level1", tape)
        t = self._stage(t, "refinement", tape)
        return engine.conv_fwd(cv["output"], t, tape, residual=img, need_res_grad=input_grad)

    # ------------------------------------------------------------------ forward
    def forward(self, img, tape: Tape | None = None, return_residual=False, input_grad=False):
        """img: [B,3,P,P] fp32 CUDA, P % 8 == 0.  Returns out (and res = img - first pass).
        input_grad: also propagate the gradient to ``img`` (training never needs it: the degraded image is data;
        the drop-in module sets it when its input requires grad)."""
        if img.dim() != 4 or img.shape[1] != 3 or img.shape[2] % 8 or img.shape[3] % 8:
            raise ValueError(f"T_net input must be [B,3,H,W] with H,W multiples of 8, got {tuple(img.shape)}")
        img = img.contiguous()
        cv = self.conv
        e1 = self._stage(engine.conv_fwd(cv["patch_embed"], img, tape, need_dx=input_grad), "encoder_level1", tape)
        e2 = self._stage(self._down(e1, "down1_2", tape), "encoder_level2", tape)
        e3 = self._stage(self._down(e2, "down2_3", tape), "encoder_level3", tape)
        l4 = self._down(e3, "down3_4", tape)
        if tape is not None:
            tape.marker(2)                 # backward: latent's weight gradients are final here
        latent = self._stage(l4, "latent", tape)
        if tape is not None:
            tape.marker(1)                 # ... the decoder side (both passes) is final here
        first = self._decode(latent, e1, e2, e3, img, tape, input_grad)
        res = engine.axpby_fwd(img, first, 1.0, -1.0, tape, need_dx=input_grad)
        if tape is not None:
            tape.marker(0)                 # ... the residual conditioner is final here
        r = self._stage(engine.conv_fwd(cv["patch_embed"], res, tape), "resencoder_level1", tape)
        r = self._stage(self._down(r, "resdown1_2", tape), "resencoder_level2", tape)
        r = self._stage(self._down(r, "resdown2_3", tape), "resencoder_level3", tape)
        r = self._stage(self._down(r, "down3_4", tape), "reslatent", tape)
        latent2 = engine.axpby_fwd(latent, r, 1.0, 0.8, tape)
        out = self._decode(latent2, e1, e2, e3, img, tape, input_grad)
        return (out, res) if return_residual else out
