"""Host-side execution engine for the transport map: hand-derived forward and backward of the
Restormer blocks and glue convolutions, expressed as sequences of ``librcot_b200`` kernel launches.

Nothing here differentiates with autograd: every op records its own backward closure on a small
tape (``Tape``), the closures call the backward kernels, and weight gradients are accumulated
straight into the parameter set's flat gradient buffer (so modules that are invoked twice per
forward -- reference Net_Restormer.py:343-432 -- sum their gradients like autograd would).

Reference behaviour followed:  TransformerBlock Net_Restormer.py:201-214, Attention :19-50,
FeedForward :67-85, LayerNorm :173-200, Downsample/Upsample/OverlapPatchEmbed :86-122,
T_net.forward :328-434.  Algebra: SURVEY.md Appendix A.
"""
from __future__ import annotations

import os

import torch

from . import ops

# RCOT_FUSED_GDFN_MID=1 selects the one-kernel form of GDFN's middle backward (csrc/dwconv.cu gdfn_mid_bwd_kernel).
# It moves 13.3 instead of 29.3 C*H*W passes but is issue-bound (8 k warp-instructions per 32x32 tile, measured
# 1.49 ms vs 1.29 ms for the two kernels at C=96, 128x128, B=32), so the two-kernel form stays the default.
FUSED_GDFN_MID = os.environ.get("RCOT_FUSED_GDFN_MID", "0") == "1"
# One-kernel GDFN forward (csrc/gdfn_fused.cu; C in {48, 96}, H % 8 == 0, W % 16 == 0) instead of the three launches
# pm_gemm -> dw_gate -> pm_gemm.  RCOT_FUSED_GDFN: "auto" (default) = wherever the hidden tensors are NOT kept for the
# backward (inference, and the recompute mode of training), "1" = always (the kernel then also writes u and g), "0" =
# never.  Measured at C=96, 128x128, B=32 (version 6 of the kernel): 0.73 ms vs 0.98 ms for the three launches with 10x
# less DRAM traffic (0.39 vs 0.50 ms at C=48), but 1.01 vs 0.98 ms when u and g must also be written -- hence "auto".
_FG = os.environ.get("RCOT_FUSED_GDFN", "auto")
FUSED_GDFN = _FG != "0"              # blobs are packed
FUSED_GDFN_ALWAYS = _FG == "1"
# One-kernel MDTA phase 1 (csrc/mdta_fused.cu: qkv GEMM -> depthwise stencil -> Gram + row norms, pre / q / k on chip)
# instead of pm_gemm -> dw_plain -> pk_gemm.  RCOT_FUSED_MDTA: "auto" (default) = wherever pre and qkv are not kept for
# the backward, "1" = always (the kernel then also writes pre, q and k), "0" = never.  Measured at C=96, 128x128, B=32:
# 0.47 vs 0.56 ms (0.23 vs 0.33 ms at C=48); 0.82 vs 0.56 ms when pre and qkv must also be written -- hence "auto".
# Direct FP32 kernels (csrc/conv3.cu) for the convs that end in three channels (output conv, data gradients of
# patch_embed and of F_net's first layer) instead of the implicit GEMM with N = 3 padded to 16: RCOT_DIRECT_CONV3=0 is
# the A/B switch back.
DIRECT_CONV3 = os.environ.get("RCOT_DIRECT_CONV3", "1") != "0"
# The matching direct WEIGHT-gradient kernel (rcot_conv3_wgrad: rolling shared-memory rows, 3k register sums per thread)
# is parity-tested but measured SLOWER than the pixel-as-K GEMMs it would replace: 6 launches = 2.66 ms against 1.81 ms
# per step at 128x128, batch 32 (96/160-thread CTAs walking a serial 128-pixel loop are latency-bound) -> opt-in.
DIRECT_WGRAD3 = os.environ.get("RCOT_DIRECT_WGRAD3", "0") == "1"
_FM = os.environ.get("RCOT_FUSED_MDTA", "auto")
FUSED_MDTA = _FM != "0"
FUSED_MDTA_ALWAYS = _FM == "1"
# RCOT_LNB_EPILOGUE=1: LayerNorm backward inside the epilogue of the GEMM that produces dL/dLN(x) (C <= 256) instead of
# its own kernel: two passes over the block tensor and 168 launches per step fewer -- but measured SLOWER (step 246.7 ->
# 265.3 ms at B=32: pm_gemm +32.8 ms, ln_bwd -13.3 ms): the four epilogue warps of a CTA become the bottleneck of the
# K-heavy dz GEMMs, while the stand-alone ln_bwd kernel spreads the same work over the whole chip.  Kept as an opt-in,
# parity-tested variant (tests/test_lnb.py).
LNB_EPILOGUE = os.environ.get("RCOT_LNB_EPILOGUE", "0") == "1"


# ---------------------------------------------------------------------------------- parameters
class ParamSet:
    """Parameters of one network as views into ONE flat fp32 buffer, a matching flat gradient
    buffer, and the packed bf16 hi/lo tcgen05 operand images of every GEMM weight."""

    def __init__(self, named_params, device, used=None, bucket_of=None):
        """named_params: ordered {name: tensor}. ``used``: names that receive gradients (the rest
        is placed at the tail of the flat buffers so optimizer kernels can skip it).
        ``bucket_of(name) -> int``: groups the used parameters into contiguous gradient buckets (bucket 0 first),
        so each bucket can be all-reduced as soon as the backward has finished it (``bucket_ranges``)."""
        names = list(named_params)
        if used is not None:
            names = [n for n in names if n in used] + [n for n in names if n not in used]
        if bucket_of is not None:
            head = [n for n in names if used is None or n in used]
            tail = [n for n in names if not (used is None or n in used)]
            names = sorted(head, key=bucket_of) + tail          # stable: state_dict order inside a bucket
        self.names = names
        # every tensor starts on a 16-byte boundary of the flat buffers (vector reductions into .grad, float4
        # optimizer passes); the padding elements stay zero in both buffers, so the optimizers leave them alone
        self.offsets = {}
        off = 0
        self.n_used = None
        for n in names:
            if used is not None and self.n_used is None and n not in used:
                self.n_used = off
            self.offsets[n] = off
            off += (named_params[n].numel() + 3) // 4 * 4
        self.numel = off
        if self.n_used is None:
            self.n_used = off
        self.bucket_ranges = [(0, self.n_used)]
        if bucket_of is not None:
            self.bucket_ranges, cur, start = [], None, 0
            for n in names:
                if used is not None and n not in used:
                    break
                b = bucket_of(n)
                if cur is not None and b != cur:
                    self.bucket_ranges.append((start, self.offsets[n]))
                    start = self.offsets[n]
                cur = b
            self.bucket_ranges.append((start, self.n_used))
        self.flat = torch.zeros(off, device=device, dtype=torch.float32)
        self.grad = torch.zeros(off, device=device, dtype=torch.float32)
        self.p, self.g = {}, {}
        for n in names:
            src = named_params[n]
            o = self.offsets[n]
            view = self.flat[o:o + src.numel()].view(src.shape)
            view.copy_(src.detach())
            self.p[n] = view
            self.g[n] = self.grad[o:o + src.numel()].view(src.shape)
        self.table = ops.PackTable(device)
        self.pack_idx = {}
        self.gdfn = {}                 # block prefix -> weight blob of the fused GDFN forward kernel
        self.gdfn_ver = {}
        self.mdta = {}                 # block prefix -> weight blob of the fused MDTA phase-1 kernel
        self.mdta_ver = {}
        self.weights_ver = 0

    def add_pack(self, name, kind):
        key = (name, kind)
        if key not in self.pack_idx:
            self.pack_idx[key] = self.table.add(self.p[name], kind)
        return key

    def finalize(self):
        self.table.finalize()
        self.repack()
        return self

    def add_gdfn(self, prefix, C, hid):
        if prefix not in self.gdfn:
            self.gdfn[prefix] = torch.empty(ops.gdfn_blob_bytes(C, hid), dtype=torch.uint8, device=self.flat.device)
            self.gdfn_ver[prefix] = -1

    def add_mdta(self, prefix, C):
        if prefix not in self.mdta:
            self.mdta[prefix] = torch.empty(ops.mdta_p1_blob_bytes(C), dtype=torch.uint8, device=self.flat.device)
            self.mdta_ver[prefix] = -1

    def mdta_blob(self, prefix):
        if self.mdta_ver[prefix] != self.weights_ver:
            a = prefix + "attn."
            ops.mdta_p1_pack(self.p[a + "qkv.weight"], self.p[a + "qkv_dwconv.weight"], self.mdta[prefix])
            self.mdta_ver[prefix] = self.weights_ver
        return self.mdta[prefix]

    def repack(self):
        if self.table.entries:
            self.table.repack()
        self.weights_ver += 1          # the fused-GDFN blobs are re-packed lazily, by the first forward that uses them

    def gdfn_blob(self, prefix):
        if self.gdfn_ver[prefix] != self.weights_ver:
            f = prefix + "ffn."
            ops.gdfn_pack(self.p[f + "project_in.weight"], self.p[f + "dwconv.weight"], self.p[f + "project_out.weight"],
                          self.gdfn[prefix])
            self.gdfn_ver[prefix] = self.weights_ver
        return self.gdfn[prefix]

    def pack(self, name, kind):
        return self.table.ptr(self.pack_idx[(name, kind)])

    def zero_grad(self):
        ops.zero_(self.grad)


# ---------------------------------------------------------------------------------- tape
class Tape:
    """Reverse-mode tape over engine ops. Gradients are keyed by tensor identity."""

    def __init__(self, enabled=True, save_hidden=False):
        self.enabled = enabled
        self.save_hidden = save_hidden   # keep the blocks' hidden tensors instead of recomputing them in backward
        self.ops = []
        self.grads = {}
        self.calls = {}
        self.on_marker = None            # callback(k): bucket k of the weight gradients is final (see marker())

    def slot(self, key):
        """How many times `key` (a block) has been invoked on this tape so far."""
        n = self.calls.get(key, 0)
        self.calls[key] = n + 1
        return n

    def record(self, out, bwd):
        if self.enabled:
            self.ops.append((out, bwd))

    def marker(self, k):
        """Placed in the FORWARD order; fires in backward once every op recorded after it has run its backward."""
        if self.enabled:
            self.ops.append((None, k))

    def add_grad(self, t, g):
        k = id(t)
        cur = self.grads.get(k)
        if cur is None:
            self.grads[k] = (t, g)
        else:
            ops.axpby(cur[1], g, 1.0, 1.0, out=cur[1])

    def backward(self, out, dout):
        """dout must be a tensor the tape may overwrite."""
        self.add_grad(out, dout)
        while self.ops:
            t, bwd = self.ops.pop()
            if t is None:
                if self.on_marker is not None:
                    self.on_marker(bwd)
                continue
            ent = self.grads.pop(id(t), None)
            if ent is not None:
                bwd(ent[1])
        leaves = {k: v for k, v in self.grads.items()}
        self.grads = {}
        return leaves

    def grad_of(self, leaves, t):
        ent = leaves.get(id(t))
        return None if ent is None else ent[1]


# ---------------------------------------------------------------------------------- scratch for MDTA small matrices
class AttnScratch:
    """Per (B, C, heads) buffers: Gram, row norms, softmax, packed per-image matrices.
    The packed buffers are zero-initialised once; the kernels only ever write the same
    head-diagonal entries, so the padding / off-diagonal zeros persist."""

    _cache = {}

    @classmethod
    def get(cls, B, C, heads, device):
        key = (B, C, heads, str(device))
        s = cls._cache.get(key)
        if s is None:
            s = cls(B, C, heads, device)
            cls._cache[key] = s
        return s

    def __init__(self, B, C, heads, device):
        c = C // heads
        self.pb = ops.packed_bytes(C, C)
        self.pb12 = ops.packed_bytes(2 * C, 2 * C)
        # one allocation for everything that must be zeroed per call: [sumsq | G | P | dA]
        n0, n1, n2 = B * 2 * C, B * (2 * C + heads * c * c), B * (2 * C + heads * c * c + C * C)
        self.zbuf = torch.zeros(n2 + B * heads * c * c, device=device)
        self.sumsq = self.zbuf[:n0].view(B, 2 * C)
        self.G = self.zbuf[n0:n1].view(B, heads, c, c)
        self.P = self.zbuf[n1:n2].view(B, C, C)
        self.dA = self.zbuf[n2:].view(B, heads, c, c)     # backward scratch: W_out^T P per head
        self.bwd_z = self.zbuf[n1:]                       # what a backward without its own forward must zero
        self.A = torch.empty(B, heads, c, c, device=device)
        self.Gt = torch.empty(B, heads, c, c, device=device)
        self.Mpack = torch.zeros(B * self.pb, dtype=torch.uint8, device=device)
        self.MTpack = torch.zeros(B * self.pb, dtype=torch.uint8, device=device)
        self.W12pack = None
        self.device = device
        self.B = B

    def w12(self):
        if self.W12pack is None:
            self.W12pack = torch.zeros(self.B * self.pb12, dtype=torch.uint8, device=self.device)
        return self.W12pack


class AttnSaved:
    """Per (block, invocation) copies of what MDTA's backward needs from its forward: row norms, softmax,
    normalised Gram and the packed M^T.  Allocated once and reused every iteration (zero-initialised pack
    padding persists because the kernels always write the same entries)."""

    def __init__(self, B, C, heads, device):
        c = C // heads
        self.sumsq = torch.zeros(B, 2 * C, device=device)
        self.A = torch.empty(B, heads, c, c, device=device)
        self.Gt = torch.empty(B, heads, c, c, device=device)
        self.MTpack = torch.zeros(B * ops.packed_bytes(C, C), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------------- Restormer block
class BlockSpec:
    """Names and sizes of one TransformerBlock inside a ParamSet."""

    def __init__(self, ps: ParamSet, prefix: str, C: int, heads: int, has_norm=True, has_attn=True, has_ffn=True):
        self.ps, self.pre, self.C, self.heads = ps, prefix, C, heads
        self.has_norm, self.has_attn, self.has_ffn = has_norm, has_attn, has_ffn
        self.saved_cache = {}
        if has_attn:
            a = prefix + "attn."
            ps.add_pack(a + "qkv.weight", "fwd")
            ps.add_pack(a + "qkv.weight", "dgrad")
            if FUSED_MDTA and C in (48, 96):
                ps.add_mdta(prefix, C)
        if has_ffn:
            f = prefix + "ffn."
            self.hid = ps.p[f + "project_out.weight"].shape[1]
            for n in ("project_in.weight", "project_out.weight"):
                ps.add_pack(f + n, "fwd")
                ps.add_pack(f + n, "dgrad")
            if FUSED_GDFN and C in (48, 96):
                ps.add_gdfn(prefix, C, self.hid)


def hidden_dtype(C, H, W):
    """Storage type of a block's hidden tensors (pre, qkv, u, g and their gradients).  bf16-storage mode
    (ops.HIDDEN_DTYPE = torch.bfloat16) applies to the levels that carry the bytes -- C <= 96 (84 % of the block traffic)
    on maps with H*W % 128 == 0 (the TMA-staged kernels) -- the deep, tensor-bound levels keep fp32."""
    if ops.HIDDEN_DTYPE == torch.bfloat16 and C <= 96 and (H * W) % 128 == 0 and W % 4 == 0:
        return torch.bfloat16
    return torch.float32


def _stats_of(x):
    """Per-pixel LayerNorm (mean, rstd) of x: left on the tensor by the GEMM epilogue that produced it
    (pm_gemm stats_out), else one ln_stats launch."""
    st = getattr(x, "_rcot_ln_stats", None)
    return st if st is not None else ops.ln_stats(x)


def _ln_args(ps, name, stats):
    return (stats, ps.p[name + ".body.weight"], ps.p[name + ".body.bias"])


def mdta_fwd(bs: BlockSpec, x, norm_name, residual, need_bwd, store=None):
    """y = [x +] project_out(attn(dwconv(qkv(LN(x))))).  Returns (y, ctx).
    ``store`` (AttnSaved) receives the small tensors the backward needs; default: the shared scratch."""
    ps, C, h = bs.ps, bs.C, bs.heads
    a = bs.pre + "attn."
    B, _, H, W = x.shape
    sc = AttnScratch.get(B, C, h, x.device)
    st = sc if store is None else store
    stats = _stats_of(x) if norm_name else None
    ln = _ln_args(ps, norm_name, stats) if norm_name else None
    c = C // h
    fused = (bs.pre in ps.mdta and (FUSED_MDTA_ALWAYS or not need_bwd) and ops.TERMS == 3
             and hidden_dtype(C, H, W) == torch.float32 and ops.mdta_p1_supported(C, H, W, h))
    if not fused:
        pre = ops.pm_gemm(x, ps.pack(a + "qkv.weight", "fwd"), 3 * C, ln=ln, out_dtype=hidden_dtype(C, H, W))
    ops.zero_(sc.zbuf)
    if store is not None:
        ops.zero_(store.sumsq)
    if fused:
        # one kernel: pre, q and k stay on chip (written out only when the backward wants them kept)
        v, pre, qkv = ops.mdta_p1(x, ps.mdta_blob(bs.pre), h, sc.G, st.sumsq, ln=ln, save=need_bwd)
        if qkv is None:
            qkv = v.new_empty(0)                       # nothing kept: only v exists
    else:
        qkv = ops.dwconv(pre, ps.p[a + "qkv_dwconv.weight"], sumsq=st.sumsq, nsq=2 * C)
        ops.pk_gemm(qkv[:, :C], qkv[:, C:2 * C], sc.G, ldo=c, per_image=True, groups=h, out_gs=c * c)
        v = qkv[:, 2 * C:]
    ops.attn_fwd(sc.G, st.sumsq, ps.p[a + "temperature"], ps.p[a + "project_out.weight"], st.A, st.Gt, sc.Mpack,
                 st.MTpack if need_bwd else None, B, C, h)
    y = ops.pm_gemm(v, sc.Mpack.data_ptr(), C, wpack_bs=sc.pb, residual=x if residual else None,
                    stats_out=bool(norm_name))      # feeds LN2 of the same block
    return y, (stats, pre, qkv, sc, st)


def mdta_bwd(bs: BlockSpec, x, dy, norm_name, residual, ctx):
    """Backward of mdta_fwd given the recomputed ctx. Returns dx (fresh tensor)."""
    ps, C, h = bs.ps, bs.C, bs.heads
    a = bs.pre + "attn."
    B, _, H, W = x.shape
    stats, pre, qkv, sc, st = ctx
    if st is not sc:
        ops.zero_(sc.bwd_z)      # [P | dA]: the forward that zeroed the shared scratch may be long gone
    ops.pk_gemm(dy, qkv[:, 2 * C:], sc.P, ldo=C, per_image=True)
    ops.attn_bwd(sc.P, st.sumsq, ps.p[a + "temperature"], ps.p[a + "project_out.weight"], st.A, st.Gt,
                 ps.g[a + "project_out.weight"], ps.g[a + "temperature"], sc.w12(), B, C, h, sc.dA)
    dqkv = torch.empty_like(qkv)
    ops.pm_gemm(qkv[:, :2 * C], sc.w12().data_ptr(), 2 * C, wpack_bs=sc.pb12, out=dqkv, out_coff=0)
    ops.pm_gemm(dy, st.MTpack.data_ptr(), C, wpack_bs=sc.pb, out=dqkv, out_coff=2 * C)
    dpre = ops.dwconv_bwd(pre, dqkv, ps.p[a + "qkv_dwconv.weight"], ps.g[a + "qkv_dwconv.weight"])
    ln = _ln_args(ps, norm_name, stats) if norm_name else None
    ops.pk_gemm(dpre, x, ps.g[a + "qkv.weight"], ldo=C, ln=ln)
    if norm_name and LNB_EPILOGUE and C <= 256:
        return ops.pm_gemm(dpre, ps.pack(a + "qkv.weight", "dgrad"), C, residual=dy if residual else None,
                           lnb=(x, stats, ps.p[norm_name + ".body.weight"], ps.g[norm_name + ".body.weight"],
                                ps.g[norm_name + ".body.bias"]))
    dz = ops.pm_gemm(dpre, ps.pack(a + "qkv.weight", "dgrad"), C, residual=None if norm_name or not residual else dy)
    if not norm_name:
        return dz
    return ops.ln_bwd(dz, x, stats, ps.p[norm_name + ".body.weight"], ps.g[norm_name + ".body.weight"],
                      ps.g[norm_name + ".body.bias"], dy=dy if residual else None, dx=dz)


def gdfn_fwd(bs: BlockSpec, x, norm_name, residual, keep=False):
    ps, C = bs.ps, bs.C
    f = bs.pre + "ffn."
    hid = bs.hid
    stats = _stats_of(x) if norm_name else None
    ln = _ln_args(ps, norm_name, stats) if norm_name else None
    hd = hidden_dtype(C, x.shape[2], x.shape[3])
    if (bs.pre in ps.gdfn and (FUSED_GDFN_ALWAYS or not keep) and ops.TERMS == 3 and hd == torch.float32
            and ops.gdfn_supported(C, x.shape[2], x.shape[3])):
        # one kernel, hidden tensor on chip; u / g are written out only when the backward wants them kept
        y, u, g = ops.gdfn_fwd(x, ps.gdfn_blob(bs.pre), hid, ln=ln, residual=residual, stats_out=bool(norm_name), save=keep)
        return (y, (stats, u, g)) if keep else y
    u = ops.pm_gemm(x, ps.pack(f + "project_in.weight", "fwd"), 2 * hid, ln=ln, out_dtype=hd)
    g = ops.dwconv(u, ps.p[f + "dwconv.weight"], mode=1)
    y = ops.pm_gemm(g, ps.pack(f + "project_out.weight", "fwd"), C, residual=x if residual else None,
                    stats_out=bool(norm_name))      # feeds LN1 of the next block
    return (y, (stats, u, g)) if keep else y


def gdfn_bwd(bs: BlockSpec, x, dy, norm_name, residual, kept=None):
    """Returns dx (fresh tensor). The hidden tensor u and the LN statistics come from ``kept`` (saved by the
    forward) or are recomputed from x."""
    ps, C = bs.ps, bs.C
    f = bs.pre + "ffn."
    hid = bs.hid
    g_kept = None
    if kept is not None:
        stats, u, g_kept = kept
    else:
        stats = _stats_of(x) if norm_name else None
    ln = _ln_args(ps, norm_name, stats) if norm_name else None
    hd = hidden_dtype(C, x.shape[2], x.shape[3])
    if kept is None:
        u = ops.pm_gemm(x, ps.pack(f + "project_in.weight", "fwd"), 2 * hid, ln=ln, out_dtype=hd)
    dg = ops.pm_gemm(dy, ps.pack(f + "project_out.weight", "dgrad"), hid, out_dtype=hd)
    g = g_kept if g_kept is not None else torch.empty_like(dg)
    if FUSED_GDFN_MID and hd == torch.float32 and ops.gdfn_mid_ok(u):
        # gate backward + transposed depthwise conv + its weight gradient in one pass over u (no [da; db] in HBM)
        du = ops.gdfn_mid_bwd(u, dg, ps.p[f + "dwconv.weight"], ps.g[f + "dwconv.weight"],
                              g_out=None if g_kept is not None else g)
        ops.pk_gemm(dy, g, ps.g[f + "project_out.weight"], ldo=hid)
        del g, dg, u
    else:
        dab = ops.dwconv(u, ps.p[f + "dwconv.weight"], mode=2, dg=dg, g_out=None if g_kept is not None else g,
                         out=torch.empty_like(u))
        ops.pk_gemm(dy, g, ps.g[f + "project_out.weight"], ldo=hid)
        del g, dg
        du = ops.dwconv_bwd(u, dab, ps.p[f + "dwconv.weight"], ps.g[f + "dwconv.weight"])
        del dab, u
    ops.pk_gemm(du, x, ps.g[f + "project_in.weight"], ldo=C, ln=ln)
    if norm_name and LNB_EPILOGUE and C <= 256:
        return ops.pm_gemm(du, ps.pack(f + "project_in.weight", "dgrad"), C, residual=dy if residual else None,
                           lnb=(x, stats, ps.p[norm_name + ".body.weight"], ps.g[norm_name + ".body.weight"],
                                ps.g[norm_name + ".body.bias"]))
    dz = ops.pm_gemm(du, ps.pack(f + "project_in.weight", "dgrad"), C,
                     residual=None if norm_name or not residual else dy)
    if not norm_name:
        return dz
    return ops.ln_bwd(dz, x, stats, ps.p[norm_name + ".body.weight"], ps.g[norm_name + ".body.weight"],
                      ps.g[norm_name + ".body.bias"], dy=dy if residual else None, dx=dz)


def block_fwd(bs: BlockSpec, x, tape: Tape | None):
    """TransformerBlock: x + MDTA(LN1(x)), then + GDFN(LN2(.)).
    Backward either recomputes the hidden tensors from the saved block input (small memory), or -- with
    tape.save_hidden -- reuses pre/qkv/u and the small attention matrices kept by the forward
    (12.3 C*H*W floats per block: 78 GB at B=32, P=128, which the 180 GB of HBM3e accommodate)."""
    if tape is not None and tape.enabled and tape.save_hidden:
        B = x.shape[0]
        key = (B, tape.slot(id(bs)))
        store = bs.saved_cache.get(key)
        if store is None:
            store = bs.saved_cache[key] = AttnSaved(B, bs.C, bs.heads, x.device)
        xm, ctx = mdta_fwd(bs, x, bs.pre + "norm1", True, True, store=store)
        y, kept = gdfn_fwd(bs, xm, bs.pre + "norm2", True, keep=True)

        def bwd_saved(dy, x=x, xm=xm, ctx=ctx, kept=kept):
            dxm = gdfn_bwd(bs, xm, dy, bs.pre + "norm2", True, kept=kept)
            tape.add_grad(x, mdta_bwd(bs, x, dxm, bs.pre + "norm1", True, ctx))
        tape.record(y, bwd_saved)
        return y
    xm, _ = mdta_fwd(bs, x, bs.pre + "norm1", True, False)
    y = gdfn_fwd(bs, xm, bs.pre + "norm2", True)
    if tape is not None and tape.enabled:
        def bwd(dy, x=x):
            xm, ctx = mdta_fwd(bs, x, bs.pre + "norm1", True, True)
            dxm = gdfn_bwd(bs, xm, dy, bs.pre + "norm2", True)
            dx = mdta_bwd(bs, x, dxm, bs.pre + "norm1", True, ctx)
            tape.add_grad(x, dx)
        tape.record(y, bwd)
    return y


# ---------------------------------------------------------------------------------- glue convolutions
class ConvSpec:
    def __init__(self, ps: ParamSet, name: str, ks: int, pad: int, need_dgrad=True):
        self.ps, self.name, self.ks, self.pad = ps, name, ks, pad
        w = ps.p[name]
        self.Cout, self.Cin = w.shape[0], w.shape[1]
        # (the two 1x1 convs that take a concatenated input keep the channel-major order)
        self.kf = ops.conv_pack_kind(w, False, concat=name.startswith("reduce_chan"))
        self.kd = ops.conv_pack_kind(w, True)
        ps.add_pack(name, self.kf)
        if need_dgrad:
            ps.add_pack(name, self.kd)


def conv_fwd(cs: ConvSpec, x, tape, x2=None, residual=None, need_dx=True, need_res_grad=True):
    """Dense conv (stride 1, no bias), optional concat input [x, x2] and residual epilogue."""
    ps = cs.ps
    direct3 = DIRECT_CONV3 and cs.ks in (3, 5) and cs.pad == cs.ks // 2 and x2 is None and ops.TERMS == 3
    if direct3 and cs.Cout == 3 and x.dtype == torch.float32:
        y = ops.conv_to3(x, ps.p[cs.name], residual=residual)           # output conv 96 -> 3: direct FP32 kernel
    elif direct3 and cs.Cin == 3 and residual is None and x.dtype == torch.float32:
        y = ops.conv_from3(x, ps.p[cs.name])                            # patch_embed 3 -> 48
    else:
        y = ops.pm_gemm(x, ps.pack(cs.name, cs.kf), cs.Cout, ks=cs.ks, pad=cs.pad, x2=x2, residual=residual,
                        tap_major=cs.kf.endswith("_tap"))
    if tape is not None and tape.enabled:
        def bwd(dy):
            Cin = cs.Cin
            w3 = DIRECT_WGRAD3 and direct3 and dy.dtype == torch.float32 and x.dtype == torch.float32 and x.shape[3] <= 256
            if w3 and cs.Cout == 3:
                ops.conv3_wgrad(x, dy, ps.g[cs.name], from3=False)      # output conv: three = dL/dy
            elif w3 and Cin == 3:
                ops.conv3_wgrad(dy, x, ps.g[cs.name], from3=True)       # patch_embed: three = the image
            else:
                ops.pk_gemm(dy, x, ps.g[cs.name].view(cs.Cout, -1), ldo=Cin * cs.ks * cs.ks, ks=cs.ks, pad=cs.pad, b2=x2)
            if need_dx:
                if direct3 and Cin == 3 and dy.dtype == torch.float32:
                    dx = ops.conv_to3(dy, ps.p[cs.name], dgrad=True)    # patch_embed's data gradient (3 channels)
                else:
                    dx = ops.pm_gemm(dy, ps.pack(cs.name, cs.kd), Cin, ks=cs.ks, pad=cs.pad, mode=1,
                                     out_hw=(x.shape[2], x.shape[3]), tap_major=cs.kd.endswith("_tap"))
                if x2 is None:
                    tape.add_grad(x, dx)
                else:
                    tape.add_grad(x, dx[:, :x.shape[1]])
                    tape.add_grad(x2, dx[:, x.shape[1]:])
            if residual is not None and need_res_grad:
                tape.add_grad(residual, dy)
        tape.record(y, bwd)
    return y


def shuffle_fwd(x, inverse, tape, out=None):
    y = ops.pixel_shuffle(x, inverse=inverse, out=out)
    if tape is not None and tape.enabled:
        def bwd(dy):
            tape.add_grad(x, ops.pixel_shuffle(dy, inverse=not inverse))
        tape.record(y, bwd)
    return y


def axpby_fwd(x, y, a, b, tape, need_dx=True):
    """z = a*x + b*y with gradients to both (x only if need_dx)."""
    z = ops.axpby(x, y, a, b)
    if tape is not None and tape.enabled:
        def bwd(dz):
            if need_dx:
                tape.add_grad(x, ops.axpby(dz, None, a, 0.0))
            tape.add_grad(y, ops.axpby(dz, None, b, 0.0))
        tape.record(z, bwd)
    return z


def shuffle_cat_fwd(t, skip, tape):
    """cat([PixelShuffle(2)(t), skip], dim=1) written into one buffer (Net_Restormer.py:368-369)."""
    B, C4, H, W = t.shape
    Cs = C4 // 4
    out = torch.empty(B, Cs + skip.shape[1], 2 * H, 2 * W, device=t.device, dtype=torch.float32)
    ops.pixel_shuffle(t, out=out[:, :Cs])
    ops.axpby(skip, None, 1.0, 0.0, out=out[:, Cs:])
    if tape is not None and tape.enabled:
        def bwd(dcat):
            tape.add_grad(t, ops.pixel_shuffle(dcat[:, :Cs], inverse=True))
            tape.add_grad(skip, dcat[:, Cs:])
        tape.record(out, bwd)
    return out


# ---------------------------------------------------------------------------------- standalone modules
class LeafProgram:
    """Program behind a stand-alone Attention / FeedForward / TransformerBlock / Downsample /
    Upsample / OverlapPatchEmbed module (the reference's classes are usable on their own)."""

    def __init__(self, named, device, kind, C, heads, strip=""):
        self.kind, self.strip = kind, strip
        self.ps = ParamSet(named, device)
        self.grad_names = {k[len(strip):] for k in named}
        if kind == "block":
            self.bs = BlockSpec(self.ps, "", C, heads)
        elif kind == "attn":
            self.bs = BlockSpec(self.ps, "", C, heads, has_ffn=False)
        elif kind == "ffn":
            self.bs = BlockSpec(self.ps, "", C, 1, has_attn=False)
        elif kind in ("down", "up", "embed"):
            self.cs = ConvSpec(self.ps, next(iter(named)), 3, 1)
        elif kind == "ln":
            pass
        else:
            raise ValueError(kind)
        self.ps.finalize()

    def pview(self, name):
        return self.ps.p[self.strip + name]

    def gview(self, name):
        return self.ps.g[self.strip + name]

    def run(self, x, tape):
        kind = self.kind
        if kind == "block":
            return block_fwd(self.bs, x, tape)
        if kind == "attn":
            bs = self.bs
            y, _ = mdta_fwd(bs, x, None, False, False)
            if tape.enabled:
                def bwd(dy):
                    _, ctx = mdta_fwd(bs, x, None, False, True)
                    tape.add_grad(x, mdta_bwd(bs, x, dy, None, False, ctx))
                tape.record(y, bwd)
            return y
        if kind == "ln":
            ps, st = self.ps, self.strip
            y, stats = ops.ln_fwd(x, ps.p[st + "body.weight"], ps.p[st + "body.bias"])
            if tape.enabled:
                tape.record(y, lambda dy: tape.add_grad(x, ops.ln_bwd(dy, x, stats, ps.p[st + "body.weight"],
                                                                      ps.g[st + "body.weight"], ps.g[st + "body.bias"])))
            return y
        if kind == "ffn":
            bs = self.bs
            y = gdfn_fwd(bs, x, None, False)
            if tape.enabled:
                tape.record(y, lambda dy: tape.add_grad(x, gdfn_bwd(bs, x, dy, None, False)))
            return y
        y = conv_fwd(self.cs, x, tape)
        if kind == "down":
            y = shuffle_fwd(y, True, tape)
        elif kind == "up":
            y = shuffle_fwd(y, False, tape)
        return y
