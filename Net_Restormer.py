"""Drop-in replacement for the reference's ``Net_Restormer.py``: same class names, constructor
signatures, attribute names (hence ``state_dict`` keys and pickled-checkpoint layout) and the
same parameter initialisation stream, but every forward/backward runs on the hand-written
sm_100a kernels of ``rcot_b200`` -- there is no PyTorch/CPU compute path behind these modules.

Reference surface mirrored (file:line in /root/reference/Net_Restormer.py):
  Attention 19-50 · FeedForward 67-85 · Downsample 86-94 · Upsample 103-111 · OverlapPatchEmbed
  113-122 · BiasFree/WithBias_LayerNorm, LayerNorm 158-200 · TransformerBlock 201-214 ·
  T_net 215-434 · F_net 436-522.

Modules hold ordinary ``nn.Parameter`` tensors.  On the first CUDA forward a module builds an
execution program (``rcot_b200.tnet.TnetProgram`` / ``rcot_b200.fnet.FnetProgram``): parameters are
re-pointed to views of one flat buffer, GEMM weights are packed for tcgen05, and autograd sees
ONE custom Function per call whose backward runs the hand-derived kernels.
"""
from __future__ import annotations

import numbers

import torch
import torch.nn as nn

SAVE_RESIDUAL_PNG = False  # the reference dumps ./checksample/res.png on EVERY forward (:433); opt-in here


def _conv(cin, cout, k, bias, stride=1, pad=None, groups=1):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=(k // 2 if pad is None else pad),
                     groups=groups, bias=bias)


# ---------------------------------------------------------------------------------- engine bridge
class _EngineModule(nn.Module):
    """Shared plumbing: lazily built program, parameter re-pointing, repack-on-change, pickling."""

    _program_factory = None  # set by subclasses: (module, named_params, device) -> program

    def _named(self):
        return {k: v for k, v in self.named_parameters()}

    def _get_program(self, device):
        named = self._named()
        prog = self.__dict__.get("_program")
        first = next(iter(named.values()))
        if prog is None or prog.pview(next(iter(named))).data_ptr() != first.data_ptr() or prog.device != device:
            prog = self._build_program(named, device)
            prog.device = device
            for k, p in named.items():          # parameters become views of the flat buffer
                p.data = prog.pview(k)
            prog.versions = None
            self.__dict__["_program"] = prog
        versions = tuple(p._version for p in named.values())
        if prog.versions != versions:            # an optimizer touched the weights -> re-pack for tcgen05
            prog.ps.repack()
            prog.versions = versions
        return prog

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_program", None)              # kernel state is rebuilt lazily; never pickled
        return state

    def _require_cuda(self, x):
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: rcot_b200 has no CPU path; move the module and input to a "
                               "B200 (`.cuda()`). The CPU oracle lives in oracle/ and is test-only.")
        if x.dtype != torch.float32:
            raise TypeError(f"{type(self).__name__}: expected float32 input, got {x.dtype}")


class _ProgramFn(torch.autograd.Function):
    """One autograd node for a whole module call; backward = hand-derived kernels via the tape."""

    @staticmethod
    def forward(ctx, prog, run, need, x, *params):
        from rcot_b200.engine import Tape
        tape = Tape(enabled=need)
        xin = x.detach().contiguous()
        out = run(prog, xin, tape)
        ctx.prog, ctx.tape, ctx.xin, ctx.out = prog, tape, xin, out
        ctx.x_needs = x.requires_grad
        return out.detach()

    @staticmethod
    @torch.autograd.function.once_differentiable     # hand-derived first-order backward: create_graph=True raises
    def backward(ctx, dout):
        prog, tape = ctx.prog, ctx.tape
        prog.ps.zero_grad()
        leaves = tape.backward(ctx.out, dout.contiguous().clone())
        dx = tape.grad_of(leaves, ctx.xin) if ctx.x_needs else None
        grads = []
        for name, need in zip(prog.param_order, ctx.needs_input_grad[4:]):
            grads.append(prog.gview(name).clone() if need and name in prog.grad_names else None)
        return (None, None, None, dx, *grads)


def _run_program(module, prog, run, x):
    named = module._named()
    prog.param_order = list(named)
    # grad mode is off inside Function.forward, so decide here whether a tape is needed
    need = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in named.values()))
    return _ProgramFn.apply(prog, run, need, x, *named.values())


# ---------------------------------------------------------------------------------- leaf modules
class Attention(_EngineModule):
    """MDTA channel attention (reference :19-50)."""

    def __init__(self, dim, num_heads, bias):
        super().__init__()
        if bias:
            raise NotImplementedError("rcot_b200 kernels implement the bias=False configuration T_net uses")
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = _conv(dim, dim * 3, 1, bias)
        self.qkv_dwconv = _conv(dim * 3, dim * 3, 3, bias, groups=dim * 3)
        self.project_out = _conv(dim, dim, 1, bias)

    def _build_program(self, named, device):
        from rcot_b200 import engine
        return engine.LeafProgram({"attn." + k: v for k, v in named.items()}, device, "attn",
                                  self.qkv.in_channels, self.num_heads, strip="attn.")

    def forward(self, x):
        self._require_cuda(x)
        prog = self._get_program(x.device)
        return _run_program(self, prog, lambda p, t, tape: p.run(t, tape), x)


class FeedForward(_EngineModule):
    """GDFN gated feed-forward (reference :67-85)."""

    def __init__(self, dim, ffn_expansion_factor, bias):
        super().__init__()
        if bias:
            raise NotImplementedError("rcot_b200 kernels implement the bias=False configuration T_net uses")
        hidden = int(dim * ffn_expansion_factor)
        self.project_in = _conv(dim, hidden * 2, 1, bias)
        self.dwconv = _conv(hidden * 2, hidden * 2, 3, bias, groups=hidden * 2)
        self.project_out = _conv(hidden, dim, 1, bias)

    def _build_program(self, named, device):
        from rcot_b200 import engine
        return engine.LeafProgram({"ffn." + k: v for k, v in named.items()}, device, "ffn",
                                  self.project_in.in_channels, 1, strip="ffn.")

    def forward(self, x):
        self._require_cuda(x)
        prog = self._get_program(x.device)
        return _run_program(self, prog, lambda p, t, tape: p.run(t, tape), x)


class BiasFree_LayerNorm(nn.Module):
    """Kept for checkpoint/class-path compatibility (reference :158-171); T_net never selects it."""

    def __init__(self, normalized_shape):
        super().__init__()
        if isinstance(normalized_shape, numbers.Integral):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = torch.Size(normalized_shape)
        assert len(self.normalized_shape) == 1
        self.weight = nn.Parameter(torch.ones(self.normalized_shape))

    def forward(self, x):
        raise NotImplementedError("BiasFree LayerNorm has no rcot_b200 kernel (unused by T_net: LayerNorm_type='WithBias')")


class WithBias_LayerNorm(nn.Module):
    """Parameter holder for the per-pixel LayerNorm over channels (reference :173-189); the
    arithmetic runs as a prologue of the GEMM that consumes it."""

    def __init__(self, normalized_shape):
        super().__init__()
        if isinstance(normalized_shape, numbers.Integral):
            normalized_shape = (normalized_shape,)
        self.normalized_shape = torch.Size(normalized_shape)
        assert len(self.normalized_shape) == 1
        self.weight = nn.Parameter(torch.ones(self.normalized_shape))
        self.bias = nn.Parameter(torch.zeros(self.normalized_shape))


class LayerNorm(_EngineModule):
    """Per-pixel LayerNorm over channels on an NCHW tensor (reference :190-200).  Inside TransformerBlock /
    T_net it runs as the prologue of the consuming GEMM; called on its own it is one `rcot_ln_fwd` launch."""

    def __init__(self, dim, LayerNorm_type):
        super().__init__()
        self.body = BiasFree_LayerNorm(dim) if LayerNorm_type == 'BiasFree' else WithBias_LayerNorm(dim)

    def _build_program(self, named, device):
        from rcot_b200 import engine
        return engine.LeafProgram(named, device, "ln", 0, 0)

    def forward(self, x):
        if isinstance(self.body, BiasFree_LayerNorm):
            return self.body(x)
        self._require_cuda(x)
        prog = self._get_program(x.device)
        return _run_program(self, prog, lambda p, t, tape: p.run(t, tape), x)


class TransformerBlock(_EngineModule):
    """x + MDTA(LN(x)), then + GDFN(LN(.)) (reference :201-214) as one engine call."""

    def __init__(self, dim, num_heads, ffn_expansion_factor, bias, LayerNorm_type):
        super().__init__()
        self.norm1 = LayerNorm(dim, LayerNorm_type)
        self.attn = Attention(dim, num_heads, bias)
        self.norm2 = LayerNorm(dim, LayerNorm_type)
        self.ffn = FeedForward(dim, ffn_expansion_factor, bias)

    def _build_program(self, named, device):
        from rcot_b200 import engine
        return engine.LeafProgram(named, device, "block", self.attn.qkv.in_channels, self.attn.num_heads)

    def forward(self, x):
        self._require_cuda(x)
        prog = self._get_program(x.device)
        return _run_program(self, prog, lambda p, t, tape: p.run(t, tape), x)


class _GlueConv(_EngineModule):
    def _build_program(self, named, device):
        from rcot_b200 import engine
        return engine.LeafProgram(named, device, self._kind, 0, 0)

    def forward(self, x):
        self._require_cuda(x)
        prog = self._get_program(x.device)
        return _run_program(self, prog, lambda p, t, tape: p.run(t, tape), x)


class Downsample(_GlueConv):
    """conv3x3 n -> n/2, PixelUnshuffle(2) (reference :86-94)."""
    _kind = "down"

    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.Sequential(_conv(n_feat, n_feat // 2, 3, False), nn.PixelUnshuffle(2))


class Upsample(_GlueConv):
    """conv3x3 n -> 2n, PixelShuffle(2) (reference :103-111)."""
    _kind = "up"

    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.Sequential(_conv(n_feat, n_feat * 2, 3, False), nn.PixelShuffle(2))


class OverlapPatchEmbed(_GlueConv):
    """conv3x3 in_c -> embed_dim (reference :113-122)."""
    _kind = "embed"

    def __init__(self, in_c=3, embed_dim=48, bias=False):
        super().__init__()
        if bias:
            raise NotImplementedError("bias=False only")
        self.proj = _conv(in_c, embed_dim, 3, bias)


# ---------------------------------------------------------------------------------- transport map
def _tnet_layout(dim, nb, nref, heads, ffn, bias, ln):
    """Construction order of the reference T_net.__init__ (:229-326): it fixes both the state_dict
    key order and the order in which the default initialisers consume the RNG."""
    d1, d2, d3, d4 = dim, dim * 2, dim * 4, dim * 8

    def stack(n, c, h):
        return lambda: nn.Sequential(*[TransformerBlock(c, h, ffn, bias, ln) for _ in range(n)])

    def block(c, h):
        return lambda: TransformerBlock(c, h, ffn, bias, ln)

    def pw(cin, cout):
        return lambda: _conv(cin, cout, 1, bias)

    return [
        ("patch_embed", lambda: OverlapPatchEmbed(3, dim)), ("res_patch_embed", lambda: OverlapPatchEmbed(3, dim)),
        ("chnl_reduce1", pw(64, 64)), ("chnl_reduce2", pw(128, 128)), ("chnl_reduce3", pw(320, 256)),
        ("reduce_noise_channel_1", pw(d1 + 64, d1)),
        ("encoder_level1", stack(nb[0], d1, heads[0])), ("resencoder_level1", stack(nb[0], d1, heads[0])),
        ("down1_2", lambda: Downsample(d1)), ("resdown1_2", lambda: Downsample(d1)),
        ("reduce_noise_channel_2", pw(d2 + 128, d2)),
        ("encoder_level2", stack(nb[1], d2, heads[1])), ("resencoder_level2", stack(nb[1], d2, heads[1])),
        ("down2_3", lambda: Downsample(d2)), ("resdown2_3", lambda: Downsample(d2)),
        ("reduce_noise_channel_3", pw(d3 + 256, d3)),
        ("encoder_level3", stack(nb[2], d3, heads[2])), ("resencoder_level3", stack(nb[2], d3, heads[2])),
        ("down3_4", lambda: Downsample(d3)), ("resdown3_4", lambda: Downsample(d3)),
        ("latent", stack(nb[3], d4, heads[3])), ("reslatent", stack(nb[3], d4, heads[3])),
        ("up4_3", lambda: Upsample(d3)), ("reduce_chan_level3", pw(d2 + 192, d3)),
        ("noise_level3", block(d3 + 192, heads[2])), ("resnoise_level3", block(d3 + 192, heads[2])),
        ("reduce_noise_level3", pw(d3 + 192, d3)), ("resreduce_noise_level3", pw(d3 + 192, d3)),
        ("decoder_level3", stack(nb[2], d3, heads[2])),
        ("up3_2", lambda: Upsample(d3)), ("reduce_chan_level2", pw(d3, d2)),
        ("noise_level2", block(d2 * 2, heads[2])), ("reduce_noise_level2", pw(d2 * 2, d3)),
        ("decoder_level2", stack(nb[1], d2, heads[1])),
        ("up2_1", lambda: Upsample(d2)),
        ("noise_level1", block(d2, heads[2])), ("reduce_noise_level1", pw(d2, d2)),
        ("decoder_level1", stack(nb[0], d2, heads[0])), ("refinement", stack(nref, d2, heads[0])),
    ]


class T_net(_EngineModule):
    """Two-pass Restormer transport map with the residual-embedding conditioner (reference :215-434)."""

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4,
                 heads=[1, 2, 4, 8], ffn_expansion_factor=2.66, bias=False, LayerNorm_type='WithBias',
                 decoder=False):
        super().__init__()
        cfg = (inp_channels, out_channels, dim, tuple(num_blocks), num_refinement_blocks, tuple(heads),
               ffn_expansion_factor, bias, LayerNorm_type)
        if cfg != (3, 3, 48, (4, 6, 6, 8), 4, (1, 2, 4, 8), 2.66, False, 'WithBias'):
            raise NotImplementedError("rcot_b200 implements the configuration trainer.py builds: "
                                      "T_net(decoder=True) with the default widths/depths/heads")
        self.decoder = decoder
        layout = _tnet_layout(dim, num_blocks, num_refinement_blocks, heads, ffn_expansion_factor, bias, LayerNorm_type)
        for i, (name, make) in enumerate(layout):
            if i == 2:
                pass  # (the reference sets self.decoder here; it is a plain attribute, no RNG use)
            setattr(self, name, make())
        self.output = _conv(dim * 2, out_channels, 3, bias)

    def _build_program(self, named, device):
        from rcot_b200.tnet import TnetProgram
        return TnetProgram(named, device)

    def forward(self, inp_img, noise_emb=None):
        if not self.decoder:
            # the reference's decoder=False path fails with a channel mismatch at up4_3 (:345-349)
            raise RuntimeError("T_net(decoder=False) is not runnable in the reference either; use decoder=True")
        self._require_cuda(inp_img)
        prog = self._get_program(inp_img.device)
        holder = {}

        want_dx = torch.is_grad_enabled() and inp_img.requires_grad

        def run(p, t, tape):
            out, res = p.forward(t, tape, return_residual=True, input_grad=want_dx)
            holder["res"] = res
            return out

        out = _run_program(self, prog, run, inp_img)
        if SAVE_RESIDUAL_PNG:
            from torchvision.utils import save_image
            save_image(holder["res"].data, './checksample/res.png')
        return out


# ---------------------------------------------------------------------------------- potential
_FNET_CONVS = ((3, 64, 5, 1, 2, True), (64, 64, 4, 2, 1, True), (64, 128, 3, 1, 1, True), (128, 128, 4, 2, 1, True),
               (128, 256, 3, 1, 1, True), (256, 256, 4, 2, 1, True), (256, 512, 3, 1, 1, False),
               (512, 512, 4, 2, 1, False), (512, 512, 3, 1, 1, False), (512, 512, 4, 2, 1, False))


class F_net(_EngineModule):
    """OT potential / critic (reference :436-522): 10 x (conv + LeakyReLU 0.2) -> fc -> fc1 ->
    LeakyReLU -> fc2; conv weights N(0, 0.02)."""

    def __init__(self, patch_size=64):
        super().__init__()
        self.patch_size = patch_size
        layers = []
        for cin, cout, k, s, p, b in _FNET_CONVS:
            layers += [_conv(cin, cout, k, b, stride=s, pad=p), nn.LeakyReLU(0.2, inplace=True)]
        self.features = nn.Sequential(*layers)
        self.LeakyReLU = nn.LeakyReLU(0.2, inplace=True)
        num_fea = int(patch_size * patch_size / 2)
        self.fc = nn.Linear(num_fea, int(num_fea / 4))
        self.fc1 = nn.Linear(int(num_fea / 4), 64)
        self.fc2 = nn.Linear(64, 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0.0, 0.02)

    def _build_program(self, named, device):
        from rcot_b200.fnet import FnetProgram
        # (a module unpickled from a reference checkpoint has no `patch_size` attribute: recover it from fc, P*P/2 inputs)
        P = self.__dict__.get("patch_size") or int(round((2 * self.fc.in_features) ** 0.5))
        return FnetProgram(named, device, P)

    def forward(self, input):
        self._require_cuda(input)
        prog = self._get_program(input.device)
        return _run_program(self, prog, lambda p, t, tape: p.forward_tape(t, tape), input)
