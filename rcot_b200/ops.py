"""Thin ctypes bindings: torch tensors -> raw pointers -> ``librcot_b200.so`` (include/rcot_b200.h).

Every function launches asynchronously on ``torch.cuda.current_stream()`` and allocates outputs
with torch's caching allocator (the library never allocates).  No op has a PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

TERMS = 3  # default GEMM precision: 3 = bf16x3 split (fp32-class), 1 = single bf16 product


class PackDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("N", C.c_int32), ("K", C.c_int32),
                ("R", C.c_int32), ("s_kouter", C.c_int32), ("s_n", C.c_int32), ("s_kinner", C.c_int32)]


class PMParams(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p), ("in2", C.c_void_p), ("in_bs", C.c_int64), ("in2_bs", C.c_int64),
        ("C1", C.c_int32), ("C2", C.c_int32), ("Hs", C.c_int32), ("Ws", C.c_int32),
        ("Hr", C.c_int32), ("Wr", C.c_int32), ("B", C.c_int32),
        ("ks", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("mode", C.c_int32),
        ("ln_stats", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("wpack", C.c_void_p), ("wpack_bs", C.c_int64), ("N", C.c_int32), ("terms", C.c_int32),
        ("out", C.c_void_p), ("out_bs", C.c_int64), ("out_coff", C.c_int32), ("act", C.c_int32),
        ("slope", C.c_float), ("accumulate", C.c_int32), ("bias", C.c_void_p),
        ("mask_y", C.c_void_p), ("mask_bs", C.c_int64), ("residual", C.c_void_p), ("res_bs", C.c_int64),
        ("stats_out", C.c_void_p), ("debug", C.c_int32), ("tap_major", C.c_int32),
        ("in_bf16", C.c_int32), ("out_bf16", C.c_int32),
        ("lnb_x", C.c_void_p), ("lnb_x_bs", C.c_int64), ("lnb_stats", C.c_void_p), ("lnb_gamma", C.c_void_p),
        ("lnb_dgamma", C.c_void_p), ("lnb_dbeta", C.c_void_p),
    ]


class PKParams(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_bs", C.c_int64), ("CA", C.c_int32),
        ("b", C.c_void_p), ("b2", C.c_void_p), ("b_bs", C.c_int64), ("b2_bs", C.c_int64),
        ("CB1", C.c_int32), ("CB2", C.c_int32),
        ("Ha", C.c_int32), ("Wa", C.c_int32), ("Hb", C.c_int32), ("Wb", C.c_int32), ("B", C.c_int32),
        ("ks", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("ln_stats", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("per_image", C.c_int32), ("terms", C.c_int32),
        ("out", C.c_void_p), ("out_bs", C.c_int64), ("ldo", C.c_int32), ("groups", C.c_int32),
        ("out_gs", C.c_int64), ("a_bf16", C.c_int32), ("b_bf16", C.c_int32),
    ]


class DWParams(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p), ("in_bs", C.c_int64), ("w", C.c_void_p), ("out", C.c_void_p), ("out_bs", C.c_int64),
        ("B", C.c_int32), ("Cn", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("mode", C.c_int32), ("flip", C.c_int32), ("hid", C.c_int32), ("nsq", C.c_int32),
        ("dg", C.c_void_p), ("dg_bs", C.c_int64), ("g_out", C.c_void_p), ("g_bs", C.c_int64),
        ("sumsq", C.c_void_p), ("bf16", C.c_int32),
    ]


class AttnParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("heads", C.c_int32), ("reserved", C.c_int32),
        ("G", C.c_void_p), ("sumsq", C.c_void_p), ("temperature", C.c_void_p), ("w_out", C.c_void_p),
        ("A", C.c_void_p), ("Gt", C.c_void_p), ("Mpack", C.c_void_p), ("MTpack", C.c_void_p),
        ("pack_bs", C.c_int64), ("P", C.c_void_p), ("dw_out", C.c_void_p), ("dtemperature", C.c_void_p),
        ("W12pack", C.c_void_p), ("pack12_bs", C.c_int64), ("dA", C.c_void_p),
    ]


_configured = False


def L():
    global _configured
    lib = _lib.lib()
    if not _configured:
        lib.rcot_packed_bytes.restype = C.c_size_t
        lib.rcot_packed_bytes.argtypes = [C.c_int, C.c_int]
        _lib.check(lib.rcot_check_device(), "check_device")
        _configured = True
    return lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t, name="tensor"):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}")
    return t


HIDDEN_DTYPE = torch.float32    # torch.bfloat16 = bf16-storage mode: the blocks' hidden tensors (pre, qkv, u, g and their
                                # gradients) are stored as bf16; block inputs/outputs, weights, statistics, accumulation
                                # and every weight gradient stay fp32 (set through rcot_b200.set_hidden_dtype)


def _act(t, name="tensor"):
    """A hidden activation: CUDA fp32 or bf16.  Returns 1 for bf16."""
    if t.dtype not in (torch.float32, torch.bfloat16) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA float32 / bfloat16 tensor, got {t.dtype} on {t.device}")
    return int(t.dtype == torch.bfloat16)


def _img_view(t, name="tensor", any_act=False):
    """NCHW tensor whose per-image block [C,H,W] is contiguous; returns batch stride in elements."""
    if any_act:
        _act(t, name)
    else:
        _f32(t, name)
    B, Cc, H, W = t.shape
    if t.stride(3) != 1 or t.stride(2) != W or t.stride(1) != H * W:
        raise ValueError(f"{name}: per-image block must be contiguous, strides={t.stride()}")
    return t.stride(0) if B > 1 else Cc * H * W


# ------------------------------------------------------------------ weight packing
_PM_DEBUG = int(os.environ.get("RCOT_PM_DEBUG", "0"))   # A/B knobs of the pm_gemm TMA variant (see gemm_pm.cu)


def packed_bytes(N: int, K: int) -> int:
    return int(L().rcot_packed_bytes(N, K))


def pack_layout(kind: str, w: torch.Tensor):
    """(N, K, R, s_kouter, s_n, s_kinner) of the [N x K] view of conv weight ``w`` [Cout, Cin, kh, kw]."""
    Cout, Cin, kh, kw = w.shape
    KK = kh * kw
    if kind == "fwd":        # B[n=co][k=(ci,ky,kx)]
        return Cout, Cin * KK, Cin * KK + 1, 0, Cin * KK, 1
    if kind == "dgrad":      # B[n=ci][k=(co,ky,kx)]
        return Cin, Cout * KK, KK, Cin * KK, KK, 1
    if kind == "fwd_tap":    # B[n=co][k=(ky,kx,ci)]
        return Cout, Cin * KK, Cin, 1, Cin * KK, KK
    if kind == "dgrad_tap":  # B[n=ci][k=(ky,kx,co)]
        return Cin, Cout * KK, Cout, 1, KK, Cin * KK
    raise ValueError(kind)


def conv_pack_kind(w: torch.Tensor, dgrad: bool, concat: bool = False) -> str:
    """Tap-major K order whenever the gathered tensor's channel count allows the fast producer path."""
    Cout, Cin, kh, kw = w.shape
    gathered = Cout if dgrad else Cin
    tap = kh * kw > 1 and gathered % 16 == 0 and not concat
    return ("dgrad" if dgrad else "fwd") + ("_tap" if tap else "")


class PackTable:
    """A set of weight tensors packed by ONE kernel launch into one flat device buffer."""

    def __init__(self, device):
        self.device = device
        self.entries = []   # (src tensor, N, K, R, s_kouter, s_n, offset)
        self.total = 0
        self.buf = None
        self.table = None
        self.max_elems = 0

    def add(self, w: torch.Tensor, kind: str) -> int:
        N, K, R, sk, sn, ski = pack_layout(kind, w)
        off = self.total
        nb = packed_bytes(N, K)
        self.entries.append((w, N, K, R, sk, sn, off, ski))
        self.total += (nb + 255) // 256 * 256
        self.max_elems = max(self.max_elems, nb // 4)
        return len(self.entries) - 1

    def finalize(self):
        self.buf = torch.empty(max(self.total, 256), dtype=torch.uint8, device=self.device)
        if not self.entries:          # a program without GEMM weights (stand-alone LayerNorm)
            return self
        arr = (PackDesc * len(self.entries))()
        for i, (w, N, K, R, sk, sn, off, ski) in enumerate(self.entries):
            arr[i] = PackDesc(w.data_ptr(), self.buf.data_ptr() + off, N, K, R, sk, sn, ski)
        raw = bytes(arr)
        self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        return self

    def ptr(self, idx: int) -> int:
        return self.buf.data_ptr() + self.entries[idx][6]

    def repack(self):
        """Re-pack every tensor from its current values (call after each optimizer step)."""
        global LAUNCHES
        LAUNCHES += 1
        e0 = e1 = None
        if PROF is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(L().rcot_pack_weights(_ptr(self.table), len(self.entries), C.c_size_t(self.max_elems), _stream()),
                   "pack_weights")
        if PROF is not None:
            e1.record()
            PROF.records.append(("pack_weights", e0, e1, self.total + sum(e[0].numel() * 4 for e in self.entries)))


def pack_single(w: torch.Tensor, kind: str):
    t = PackTable(w.device)
    t.add(w, kind)
    t.finalize().repack()
    return t


# ------------------------------------------------------------------ pixel-as-M GEMM
def pm_gemm(x, wpack_ptr, N, *, ks=1, stride=1, pad=0, mode=0, x2=None, out=None, out_hw=None, out_coff=0,
            ln=None, bias=None, act=False, slope=0.2, mask_y=None, residual=None, accumulate=False,
            wpack_bs=0, terms=None, debug=0, tap_major=False, stats_out=False, out_dtype=torch.float32, lnb=None):
    """out[b, coff+n, p] = epi(sum_k A(b,p,k) W[n,k]).  ``ln`` = (stats[B,HW,2], gamma, beta).
    x may be a bf16 tensor and/or out_dtype (or ``out``) bf16: bf16-storage mode of the hidden tensors (1x1 only).
    ``lnb`` = (x_ln, stats, gamma, dgamma, dbeta): LayerNorm-backward epilogue -- the GEMM result is dL/dLN(x_ln) and
    out = [residual +] LN'(.) while dgamma / dbeta are accumulated (N <= 256)."""
    B, C1, Hs, Ws = x.shape
    in_bs = _img_view(x, "x", True)
    C2 = 0 if x2 is None else x2.shape[1]
    in2_bs = 0 if x2 is None else _img_view(x2, "x2")
    if out_hw is None:
        if mode == 0:
            out_hw = ((Hs + 2 * pad - ks) // stride + 1, (Ws + 2 * pad - ks) // stride + 1)
        else:
            raise ValueError("dgrad needs out_hw")
    Hr, Wr = out_hw
    if out is None:
        out = torch.empty(B, N, Hr, Wr, device=x.device, dtype=out_dtype)
    out_bs = _img_view(out, "out", True)
    p = PMParams()
    p.in_bf16, p.out_bf16 = _act(x, "x"), _act(out, "out")
    p.in_, p.in2, p.in_bs, p.in2_bs = x.data_ptr(), (None if x2 is None else x2.data_ptr()), in_bs, in2_bs
    p.C1, p.C2, p.Hs, p.Ws, p.Hr, p.Wr, p.B = C1, C2, Hs, Ws, Hr, Wr, B
    p.ks, p.stride, p.pad, p.mode = ks, stride, pad, mode
    if ln is not None:
        stats, gamma, beta = ln
        p.ln_stats, p.ln_gamma, p.ln_beta = _f32(stats).data_ptr(), _f32(gamma).data_ptr(), _f32(beta).data_ptr()
    p.wpack, p.wpack_bs, p.N, p.terms = wpack_ptr, wpack_bs, N, (TERMS if terms is None else terms)
    p.out, p.out_bs, p.out_coff = out.data_ptr(), out_bs, out_coff
    p.act, p.slope, p.accumulate = int(act), slope, int(accumulate)
    p.debug = debug
    p.tap_major = int(tap_major)
    p.debug = debug or _PM_DEBUG
    st = None
    if stats_out and N <= 256 and out_coff == 0:
        # LayerNorm statistics of the output rows come out of the epilogue; the consumer finds them on the tensor
        st = torch.empty(B, Hr * Wr, 2, device=x.device, dtype=torch.float32)
        p.stats_out = st.data_ptr()
    p.bias = None if bias is None else _f32(bias).data_ptr()
    if lnb is not None:
        xl, stl, gl, dgl, dbl = lnb
        p.lnb_x, p.lnb_x_bs, p.lnb_stats = xl.data_ptr(), _img_view(xl, "lnb x"), _f32(stl).data_ptr()
        p.lnb_gamma, p.lnb_dgamma, p.lnb_dbeta = _f32(gl).data_ptr(), _f32(dgl).data_ptr(), _f32(dbl).data_ptr()
    if mask_y is not None:
        p.mask_y, p.mask_bs = mask_y.data_ptr(), _img_view(mask_y, "mask_y")
    if residual is not None:
        p.residual, p.res_bs = residual.data_ptr(), _img_view(residual, "residual")
    _lib.check(L().rcot_pm_gemm(C.byref(p), _stream()), "pm_gemm")
    if st is not None:
        out._rcot_ln_stats = st
    return out


# ------------------------------------------------------------------ pixel-as-K GEMM
def pk_gemm(a, b, out, *, ldo, ks=1, stride=1, pad=0, b2=None, ln=None, per_image=False, groups=1, out_gs=0,
            CA=None, CB=None, terms=None):
    """out[(b,g,) m, n] += sum_q a[b, g*CA+m, q] * Bg(b, g, n, q); ``out`` is accumulated (atomics)."""
    B, CAf, Ha, Wa = a.shape
    _, CBf, Hb, Wb = b.shape
    p = PKParams()
    p.a, p.a_bs, p.CA = a.data_ptr(), _img_view(a, "a", True), (CAf // groups if CA is None else CA)
    p.b, p.b_bs, p.CB1 = b.data_ptr(), _img_view(b, "b", True), (CBf // groups if CB is None else CB)
    p.a_bf16, p.b_bf16 = _act(a, "a"), _act(b, "b")
    if b2 is not None:
        p.b2, p.b2_bs, p.CB2 = b2.data_ptr(), _img_view(b2, "b2"), b2.shape[1]
    p.Ha, p.Wa, p.Hb, p.Wb, p.B = Ha, Wa, Hb, Wb, B
    p.ks, p.stride, p.pad = ks, stride, pad
    if ln is not None:
        stats, gamma, beta = ln
        p.ln_stats, p.ln_gamma, p.ln_beta = stats.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    p.per_image, p.terms = int(per_image), (TERMS if terms is None else terms)
    _f32(out, "out")
    p.out, p.ldo, p.groups, p.out_gs = out.data_ptr(), ldo, groups, out_gs
    p.out_bs = out.stride(0) if per_image else 0
    _lib.check(L().rcot_pk_gemm(C.byref(p), _stream()), "pk_gemm")
    return out


# ------------------------------------------------------------------ fused GDFN forward
class GdfnParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_bs", C.c_int64), ("ln_stats", C.c_void_p), ("ln_gamma", C.c_void_p),
                ("ln_beta", C.c_void_p), ("wblob", C.c_void_p), ("y", C.c_void_p), ("y_bs", C.c_int64),
                ("stats_out", C.c_void_p), ("save_u", C.c_void_p), ("u_bs", C.c_int64), ("save_g", C.c_void_p),
                ("g_bs", C.c_int64), ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("hid", C.c_int32), ("residual", C.c_int32), ("debug", C.c_int32)]


def gdfn_supported(Cc, H, W):
    return bool(L().rcot_gdfn_supported(Cc, H, W))


def gdfn_blob_bytes(Cc, hid):
    L().rcot_gdfn_blob_bytes.restype = C.c_size_t
    return int(L().rcot_gdfn_blob_bytes(Cc, hid))


def gdfn_pack(w_in, w_dw, w_out, blob):
    """Weight blob of the fused GDFN kernel from project_in [2hid,C,1,1], dwconv [2hid,1,3,3], project_out [C,hid,1,1]."""
    global LAUNCHES
    LAUNCHES += 1
    Cc, hid = w_out.shape[0], w_out.shape[1]
    _lib.check(L().rcot_gdfn_pack(_ptr(_f32(w_in)), _ptr(_f32(w_dw)), _ptr(_f32(w_out)), _ptr(blob), Cc, hid, _stream()),
               "gdfn_pack")
    return blob


def gdfn_fwd(x, blob, hid, ln=None, residual=True, stats_out=False, save=False):
    """One-kernel GDFN forward.  Returns (y, u or None, g or None); LayerNorm statistics of y are left on
    ``y._rcot_ln_stats`` when stats_out."""
    B, Cc, H, W = x.shape
    p = GdfnParams()
    p.x, p.x_bs = x.data_ptr(), _img_view(x, "x")
    if ln is not None:
        stats, gamma, beta = ln
        p.ln_stats, p.ln_gamma, p.ln_beta = stats.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    p.wblob = blob.data_ptr()
    y = torch.empty_like(x)
    p.y, p.y_bs = y.data_ptr(), _img_view(y, "y")
    if stats_out:
        st = torch.empty(B, H * W, 2, device=x.device, dtype=torch.float32)
        p.stats_out = st.data_ptr()
        y._rcot_ln_stats = st
    u = g = None
    if save:
        u = torch.empty(B, 2 * hid, H, W, device=x.device, dtype=torch.float32)
        g = torch.empty(B, hid, H, W, device=x.device, dtype=torch.float32)
        p.save_u, p.u_bs, p.save_g, p.g_bs = u.data_ptr(), _img_view(u, "u"), g.data_ptr(), _img_view(g, "g")
    p.B, p.C, p.H, p.W, p.hid, p.residual = B, Cc, H, W, hid, int(bool(residual))
    p.debug = int(os.environ.get("RCOT_GDFN_DEBUG", "0"))
    _lib.check(L().rcot_gdfn_fwd(C.byref(p), _stream()), "gdfn_fwd")
    return y, u, g


def conv_to3(x, weight, dgrad=False, residual=None):
    """Direct kernel for a stride-1 'same' conv that ends in 3 channels: forward with ``weight`` [3, Cin, k, k], or
    (dgrad) the data gradient of a 3 -> Cout conv given dL/dy and its ``weight`` [Cout, 3, k, k].  k in {3, 5}."""
    B, Cin, H, W = x.shape
    ks = weight.shape[2]
    if (weight.shape[1] if dgrad else weight.shape[0]) != 3 or (weight.shape[0] if dgrad else weight.shape[1]) != Cin:
        raise ValueError(f"conv_to3: weight {tuple(weight.shape)} does not match x {tuple(x.shape)} (dgrad={dgrad})")
    out = torch.empty(B, 3, H, W, device=x.device, dtype=torch.float32)
    _lib.check(L().rcot_conv_to3(_ptr(x), C.c_int64(_img_view(x, "x")), _ptr(_f32(weight)), int(bool(dgrad)), _ptr(out),
                                 C.c_int64(3 * H * W), _ptr(residual),
                                 C.c_int64(_img_view(residual, "residual") if residual is not None else 0),
                                 B, Cin, H, W, ks, _stream()), "conv_to3")
    return out


def conv_from3(x, weight, bias=None, act=False, slope=0.2, mask_y=None):
    """Direct kernel for a stride-1 'same' conv that starts from 3 channels (``weight`` [Cout, 3, k, k], k in {3, 5}):
    optional bias + LeakyReLU, or the LeakyReLU-derivative mask of a previous forward (tangent pass)."""
    B, Cin, H, W = x.shape
    Cout, ks = weight.shape[0], weight.shape[2]
    if Cin != 3 or weight.shape[1] != 3:
        raise ValueError(f"conv_from3: weight {tuple(weight.shape)} / x {tuple(x.shape)} are not a 3-channel conv")
    out = torch.empty(B, Cout, H, W, device=x.device, dtype=torch.float32)
    _lib.check(L().rcot_conv_from3(_ptr(x), C.c_int64(_img_view(x, "x")), _ptr(_f32(weight)), _ptr(bias), _ptr(out),
                                   C.c_int64(Cout * H * W), _ptr(mask_y),
                                   C.c_int64(_img_view(mask_y, "mask_y") if mask_y is not None else 0), int(bool(act)),
                                   C.c_float(slope), B, Cout, H, W, ks, _stream()), "conv_from3")
    return out


def conv3_wgrad(many, three, dw, from3):
    """dw += weight gradient of a stride-1 'same' conv with 3 channels on one side.  ``from3``: the conv maps 3 -> Cm
    channels (dw [Cm, 3, k, k]; many = dL/dy, three = the conv's input); else Cm -> 3 (dw [3, Cm, k, k]; many = the conv's
    input, three = dL/dy)."""
    B, Cm, H, W = many.shape
    ks = dw.shape[2]
    want = (Cm, 3, ks, ks) if from3 else (3, Cm, ks, ks)
    if tuple(dw.shape) != want or tuple(three.shape) != (B, 3, H, W):
        raise ValueError(f"conv3_wgrad: dw {tuple(dw.shape)} / three {tuple(three.shape)} do not match many {tuple(many.shape)}")
    _lib.check(L().rcot_conv3_wgrad(_ptr(many), C.c_int64(_img_view(many, "many")), _ptr(three),
                                    C.c_int64(_img_view(three, "three")), _ptr(_f32(dw)), int(bool(from3)), B, Cm, H, W, ks,
                                    _stream()), "conv3_wgrad")
    return dw


class MdtaP1Params(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_bs", C.c_int64), ("ln_stats", C.c_void_p), ("ln_gamma", C.c_void_p),
                ("ln_beta", C.c_void_p), ("wblob", C.c_void_p), ("v", C.c_void_p), ("v_bs", C.c_int64),
                ("G", C.c_void_p), ("sumsq", C.c_void_p), ("save_pre", C.c_void_p), ("pre_bs", C.c_int64),
                ("save_qk", C.c_void_p), ("qk_bs", C.c_int64), ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("heads", C.c_int32), ("debug", C.c_int32)]


def mdta_p1_supported(Cc, H, W, heads):
    return bool(L().rcot_mdta_p1_supported(Cc, H, W, heads))


def mdta_p1_blob_bytes(Cc):
    L().rcot_mdta_p1_blob_bytes.restype = C.c_size_t
    return int(L().rcot_mdta_p1_blob_bytes(Cc))


def mdta_p1_pack(w_qkv, w_dw, blob):
    """Weight blob of the fused MDTA phase-1 kernel from qkv.weight [3C,C,1,1] and qkv_dwconv.weight [3C,1,3,3]."""
    global LAUNCHES
    LAUNCHES += 1
    Cc = w_qkv.shape[1]
    _lib.check(L().rcot_mdta_p1_pack(_ptr(_f32(w_qkv)), _ptr(_f32(w_dw)), _ptr(blob), Cc, _stream()), "mdta_p1_pack")
    return blob


def mdta_p1(x, blob, heads, G, sumsq, ln=None, save=False):
    """One-kernel MDTA phase 1: accumulates the per-head Grams into ``G`` [B, heads, c, c] and the row sums of squares
    into ``sumsq`` [B, 2C] (both zeroed by the caller) and returns (v, pre, qkv): ``v`` [B, C, H, W] alone when
    ``save`` is false (pre = qkv = None), else a view of the v part of the saved ``qkv`` tensor."""
    B, Cc, H, W = x.shape
    p = MdtaP1Params()
    p.x, p.x_bs = x.data_ptr(), _img_view(x, "x")
    if ln is not None:
        stats, gamma, beta = ln
        p.ln_stats, p.ln_gamma, p.ln_beta = stats.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    p.wblob = blob.data_ptr()
    pre = qkv = None
    if save:
        pre = torch.empty(B, 3 * Cc, H, W, device=x.device, dtype=torch.float32)
        qkv = torch.empty(B, 3 * Cc, H, W, device=x.device, dtype=torch.float32)
        v = qkv[:, 2 * Cc:]
        p.save_pre, p.pre_bs = pre.data_ptr(), _img_view(pre, "pre")
        p.save_qk, p.qk_bs = qkv.data_ptr(), 3 * Cc * H * W
        p.v, p.v_bs = v.data_ptr(), 3 * Cc * H * W
    else:
        v = torch.empty(B, Cc, H, W, device=x.device, dtype=torch.float32)
        p.v, p.v_bs = v.data_ptr(), _img_view(v, "v")
    p.G, p.sumsq = G.data_ptr(), sumsq.data_ptr()
    p.B, p.C, p.H, p.W, p.heads = B, Cc, H, W, heads
    p.debug = int(os.environ.get("RCOT_MDTA_DEBUG", "0"))
    _lib.check(L().rcot_mdta_p1(C.byref(p), _stream()), "mdta_p1")
    return v, pre, qkv


def gdfn_profile_read():
    """Cycle counters of CTA 0 of the last gdfn_fwd launched with RCOT_GDFN_DEBUG & 16: [22 warps][8] (measurement aid)."""
    import numpy as np
    out = np.zeros(22 * 8, dtype=np.uint64)
    _lib.check(L().rcot_gdfn_profile_read(out.ctypes.data_as(C.c_void_p), out.size), "gdfn_profile_read")
    return out.reshape(22, 8)


# ------------------------------------------------------------------ LayerNorm
def ln_stats(x, out=None):
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty(B, H * W, 2, device=x.device, dtype=torch.float32)
    _lib.check(L().rcot_ln_stats(_ptr(x), C.c_int64(_img_view(x, "x")), B, Cc, H * W, _ptr(out), _stream()), "ln_stats")
    return out


def ln_fwd(x, gamma, beta):
    """Stand-alone LayerNorm forward; returns (y, stats)."""
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    stats = torch.empty(B, H * W, 2, device=x.device, dtype=torch.float32)
    _lib.check(L().rcot_ln_fwd(_ptr(x), C.c_int64(_img_view(x, "x")), _ptr(_f32(gamma)), _ptr(_f32(beta)), _ptr(y),
                               C.c_int64(_img_view(y, "y")), B, Cc, H * W, _ptr(stats), _stream()), "ln_fwd")
    return y, stats


def ln_bwd(dz, x, stats, gamma, dgamma, dbeta, dy=None, dx=None):
    B, Cc, H, W = x.shape
    if dx is None:
        dx = torch.empty_like(x)
    _lib.check(L().rcot_ln_bwd(_ptr(dz), C.c_int64(_img_view(dz, "dz")), _ptr(x), C.c_int64(_img_view(x, "x")),
                               _ptr(stats), _ptr(gamma), _ptr(dy), C.c_int64(0 if dy is None else _img_view(dy, "dy")),
                               _ptr(dx), C.c_int64(_img_view(dx, "dx")), _ptr(dgamma), _ptr(dbeta), B, Cc, H * W,
                               _stream()), "ln_bwd")
    return dx


# ------------------------------------------------------------------ depthwise 3x3
def dwconv(x, w, *, out=None, mode=0, flip=False, dg=None, g_out=None, sumsq=None, nsq=0):
    """x (and out / dg / g_out) may be bf16 tensors (bf16-storage mode): same kernels, 2-byte loads and stores."""
    B, Cn, H, W = x.shape
    hid = Cn // 2 if mode != 0 else 0
    bf = _act(x, "x")
    if out is None:
        out = torch.empty(B, hid if mode == 1 else Cn, H, W, device=x.device, dtype=x.dtype)
    for t, nm in ((out, "out"), (dg, "dg"), (g_out, "g_out")):
        if t is not None and t.dtype != x.dtype:
            raise TypeError(f"dwconv: {nm} is {t.dtype} but x is {x.dtype}")
    p = DWParams()
    p.in_, p.in_bs, p.w, p.out, p.out_bs = (x.data_ptr(), _img_view(x, "x", True), _f32(w).data_ptr(), out.data_ptr(),
                                            _img_view(out, "out", True))
    p.B, p.Cn, p.H, p.W = B, Cn, H, W
    p.mode, p.flip, p.hid, p.nsq, p.bf16 = mode, int(flip), hid, nsq, bf
    if dg is not None:
        p.dg, p.dg_bs = dg.data_ptr(), _img_view(dg, "dg", True)
    if g_out is not None:
        p.g_out, p.g_bs = g_out.data_ptr(), _img_view(g_out, "g_out", True)
    if sumsq is not None:
        p.sumsq = sumsq.data_ptr()
    _lib.check(L().rcot_dwconv3x3(C.byref(p), _stream()), "dwconv3x3")
    return out


def dwconv_wgrad(x, dout, dw):
    B, Cn, H, W = x.shape
    _lib.check(L().rcot_dwconv3x3_wgrad(_ptr(x), C.c_int64(_img_view(x, "x")), _ptr(dout),
                                        C.c_int64(_img_view(dout, "dout")), _ptr(_f32(dw)), B, Cn, H, W, _stream()),
               "dwconv3x3_wgrad")
    return dw


def dwconv_bwd(x, dout, w, dw):
    """din = dw^T(dout); dw += corr(x, dout)  (x = the depthwise conv's forward input).  x / dout fp32 or both bf16."""
    B, Cn, H, W = x.shape
    bf = _act(x, "x")
    if dout.dtype != x.dtype:
        raise TypeError(f"dwconv_bwd: dout is {dout.dtype} but x is {x.dtype}")
    din = torch.empty_like(dout)
    _lib.check(L().rcot_dwconv3x3_bwd_t(_ptr(x), C.c_int64(_img_view(x, "x", True)), _ptr(dout),
                                        C.c_int64(_img_view(dout, "dout", True)), _ptr(_f32(w)), _ptr(din),
                                        C.c_int64(_img_view(din, "din", True)), _ptr(_f32(dw)), B, Cn, H, W, bf, _stream()),
               "dwconv3x3_bwd")
    return din


def gdfn_mid_ok(u):
    """Geometry the fused GDFN middle backward accepts (every level of a training patch whose width is a multiple of 32)."""
    return u.shape[3] % 32 == 0 and u.shape[2] % 4 == 0 and u.data_ptr() % 16 == 0 and _img_view(u, "u") % 4 == 0 and u.shape[1] % 2 == 0


def gdfn_mid_bwd(u, dg, w, dw, g_out=None):
    """du = dw^T(gate'(dw(u)) * dg), dw += corr(u, .), optionally g_out = gelu(a)*b: one pass (csrc/dwconv.cu)."""
    B, Cn, H, W = u.shape
    hid = Cn // 2
    du = torch.empty_like(u)
    _lib.check(L().rcot_gdfn_mid_bwd(_ptr(u), C.c_int64(_img_view(u, "u")), _ptr(dg), C.c_int64(_img_view(dg, "dg")),
                                     _ptr(_f32(w)), _ptr(du), C.c_int64(_img_view(du, "du")), _ptr(_f32(dw)),
                                     _ptr(g_out), C.c_int64(_img_view(g_out, "g_out") if g_out is not None else 0),
                                     B, hid, H, W, _stream()), "gdfn_mid_bwd")
    return du


# ------------------------------------------------------------------ MDTA small-matrix steps
def attn_fwd(G, sumsq, temperature, w_out, A, Gt, Mpack, MTpack, B, Cc, heads):
    p = AttnParams()
    p.B, p.C, p.heads = B, Cc, heads
    p.G, p.sumsq, p.temperature, p.w_out = G.data_ptr(), sumsq.data_ptr(), temperature.data_ptr(), w_out.data_ptr()
    p.A, p.Gt, p.Mpack = A.data_ptr(), Gt.data_ptr(), Mpack.data_ptr()
    p.MTpack = None if MTpack is None else MTpack.data_ptr()
    p.pack_bs = packed_bytes(Cc, Cc)
    _lib.check(L().rcot_attn_fwd(C.byref(p), _stream()), "attn_fwd")


def attn_bwd(P, sumsq, temperature, w_out, A, Gt, dw_out, dtemp, W12pack, B, Cc, heads, dA):
    """``dA``: zeroed [B, heads, c, c] scratch (phase 1 accumulates W_out^T P into it, phase 2 consumes it)."""
    p = AttnParams()
    p.dA = dA.data_ptr()
    p.B, p.C, p.heads = B, Cc, heads
    p.sumsq, p.temperature, p.w_out = sumsq.data_ptr(), temperature.data_ptr(), w_out.data_ptr()
    p.A, p.Gt, p.P = A.data_ptr(), Gt.data_ptr(), P.data_ptr()
    p.dw_out, p.dtemperature, p.W12pack = dw_out.data_ptr(), dtemp.data_ptr(), W12pack.data_ptr()
    p.pack12_bs = packed_bytes(2 * Cc, 2 * Cc)
    _lib.check(L().rcot_attn_bwd(C.byref(p), _stream()), "attn_bwd")


# ------------------------------------------------------------------ data movement
def pixel_shuffle(x, inverse=False, out=None):
    B, Cc, H, W = x.shape
    if inverse:   # [C,2H,2W] -> [4C,H,W]
        Cs, Hs, Ws = Cc, H // 2, W // 2
        shape = (B, 4 * Cc, Hs, Ws)
    else:         # [4C,H,W] -> [C,2H,2W]
        Cs, Hs, Ws = Cc // 4, H, W
        shape = (B, Cs, 2 * H, 2 * W)
    if out is None:
        out = torch.empty(shape, device=x.device, dtype=torch.float32)
    _lib.check(L().rcot_pixel_shuffle(_ptr(x), C.c_int64(_img_view(x, "x")), _ptr(out), C.c_int64(_img_view(out, "out")),
                                      B, Cs, Hs, Ws, int(inverse), _stream()), "pixel_shuffle")
    return out


def axpby(x, y=None, a=1.0, b=1.0, a_vec=None, out=None):
    """out = a*x + b*y per image block; with a_vec[B]: out = a_vec*x + (1-a_vec)*y."""
    B = x.shape[0]
    n = x[0].numel()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    _lib.check(L().rcot_axpby(_ptr(out), C.c_int64(_img_view(out, "out")), _ptr(x), C.c_int64(_img_view(x, "x")),
                              _ptr(y), C.c_int64(0 if y is None else _img_view(y, "y")), C.c_float(a), C.c_float(b),
                              _ptr(a_vec), B, C.c_int64(n), _stream()), "axpby")
    return out


def channel_sum(x, out):
    B, Cc, H, W = x.shape
    _lib.check(L().rcot_channel_sum(_ptr(x), C.c_int64(_img_view(x, "x")), _ptr(out), B, Cc, H * W, _stream()),
               "channel_sum")
    return out


def zero_(t):
    _lib.check(L().rcot_zero(_ptr(t), C.c_size_t(t.numel() * t.element_size()), _stream()), "zero")
    return t


# ------------------------------------------------------------------ F_net fully connected tail
def linear_fwd(x, W, bias=None, act=False, mask=None, slope=0.2):
    B, K = x.shape
    O = W.shape[0]
    y = torch.empty(B, O, device=x.device, dtype=torch.float32)
    _lib.check(L().rcot_linear_fwd(_ptr(x), _ptr(W), _ptr(bias), _ptr(mask), _ptr(y), B, K, O, int(act),
                                   C.c_float(slope), _stream()), "linear_fwd")
    return y


def linear_dgrad(dy, W, mask=None, slope=0.2):
    B, O = dy.shape
    K = W.shape[1]
    dx = torch.empty(B, K, device=dy.device, dtype=torch.float32)
    _lib.check(L().rcot_linear_dgrad(_ptr(dy), _ptr(W), _ptr(mask), _ptr(dx), B, K, O, C.c_float(slope), _stream()),
               "linear_dgrad")
    return dx


def linear_wgrad(dy, x, dW, db=None):
    B, O = dy.shape
    K = x.shape[1]
    _lib.check(L().rcot_linear_wgrad(_ptr(dy), _ptr(x), _ptr(dW), _ptr(db), B, K, O, _stream()), "linear_wgrad")


# ------------------------------------------------------------------ objective
def cost_stage1(out, degraded, target, de_id, gfou, acc):
    B, _, P, _ = out.shape
    _lib.check(L().rcot_cost_stage1(_ptr(out), _ptr(degraded), _ptr(target), _ptr(de_id), _ptr(gfou), _ptr(acc), B, P,
                                    _stream()), "cost_stage1")


def cost_stage2(out, degraded, target, gfou, dF, acc, dout, sigma, Sigma, n_global):
    _lib.check(L().rcot_cost_stage2(_ptr(out), _ptr(degraded), _ptr(target), _ptr(gfou), _ptr(dF), _ptr(acc), _ptr(dout),
                                    C.c_float(sigma), C.c_float(Sigma), C.c_double(n_global), C.c_int64(out.numel()),
                                    _stream()), "cost_stage2")


def sample_sumsq(x, out):
    B = x.shape[0]
    _lib.check(L().rcot_sample_sumsq(_ptr(x), _ptr(out), B, C.c_int64(x[0].numel()), _stream()), "sample_sumsq")


def gp_coef(sumsq, coef, loss, B_global):
    _lib.check(L().rcot_gp_coef(_ptr(sumsq), _ptr(coef), _ptr(loss), sumsq.numel(), B_global, _stream()), "gp_coef")


def signed_sum(x, out, n_neg, scale):
    _lib.check(L().rcot_signed_sum(_ptr(x), _ptr(out), x.numel(), n_neg, C.c_float(scale), _stream()), "signed_sum")


def rmsprop(p, g, sq, n, lr, alpha=0.99, eps=1e-8, gscale=1.0):
    _lib.check(L().rcot_rmsprop(_ptr(p), _ptr(g), _ptr(sq), C.c_int64(n), C.c_float(lr), C.c_float(alpha),
                                C.c_float(eps), C.c_float(gscale), _stream()), "rmsprop")


def rmsprop_h(p, g, sq, n, hyper, lr_mult=1.0, alpha=0.99, eps=1e-8, gscale=1.0):
    _lib.check(L().rcot_rmsprop_h(_ptr(p), _ptr(g), _ptr(sq), C.c_int64(n), _ptr(hyper), C.c_float(lr_mult),
                                  C.c_float(alpha), C.c_float(eps), C.c_float(gscale), _stream()), "rmsprop_h")


def adam_h(p, g, m, v, n, hyper, lr_mult=1.0, b1=0.9, b2=0.999, eps=1e-8, gscale=1.0):
    _lib.check(L().rcot_adam_h(_ptr(p), _ptr(g), _ptr(m), _ptr(v), C.c_int64(n), _ptr(hyper), C.c_float(lr_mult),
                               C.c_float(b1), C.c_float(b2), C.c_float(eps), C.c_float(gscale), _stream()), "adam_h")


def adam(p, g, m, v, n, lr, step, b1=0.9, b2=0.999, eps=1e-8, gscale=1.0):
    _lib.check(L().rcot_adam(_ptr(p), _ptr(g), _ptr(m), _ptr(v), C.c_int64(n), C.c_float(lr), C.c_float(b1),
                             C.c_float(b2), C.c_float(eps), step, C.c_float(gscale), _stream()), "adam")


# ------------------------------------------------------------------ launch accounting / per-kernel timing
class Profiler:
    """Per-op CUDA-event timing on the launching stream plus the op's ALGORITHMIC bytes (every
    operand counted once: what an ideal kernel must move), for bench.py's roofline block."""

    def __init__(self, detail=False):
        self.records = []
        self.detail = detail        # key ops by shape as well (offline analysis)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, nbytes in self.records:
            a = agg.setdefault(name, [0, 0.0, 0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += nbytes
        return {k: {"launches": v[0], "ms": v[1], "bytes": v[2]} for k, v in agg.items()}


PROF: Profiler | None = None
LAUNCHES = 0


def _nb(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


def _instrument(name, bytes_fn):
    def deco(fn):
        def wrapped(*a, **k):
            global LAUNCHES
            LAUNCHES += 1
            if PROF is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            key = name
            if PROF.detail:
                shp = [tuple(t.shape) for t in a if isinstance(t, torch.Tensor)][:2]
                key = f"{name} {shp} N={a[2] if name == 'pm_gemm' else ''} ks={k.get('ks', 1)} mode={k.get('mode', 0)}"
            PROF.records.append((key, e0, e1, int(bytes_fn(a, k, r))))
            return r
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def _pm_bytes(a, k, r):
    x, N = a[0], a[2]
    ks = k.get("ks", 1)
    Kdim = (x.shape[1] + (k["x2"].shape[1] if k.get("x2") is not None else 0)) * ks * ks
    return (_nb(x, k.get("x2"), k.get("residual"), k.get("mask_y"), k["lnb"][0] if k.get("lnb") else None) +
            r.shape[0] * N * r.shape[2] * r.shape[3] * r.element_size() +
            N * Kdim * 4 * (r.shape[0] if k.get("wpack_bs", 0) else 1))


pm_gemm = _instrument("pm_gemm", _pm_bytes)(pm_gemm)
pk_gemm = _instrument("pk_gemm", lambda a, k, r: _nb(a[0], a[1], k.get("b2"), a[2]))(pk_gemm)
gdfn_fwd = _instrument("gdfn_fwd", lambda a, k, r: _nb(a[0], r[0], r[1], r[2]))(gdfn_fwd)
mdta_p1 = _instrument("mdta_p1", lambda a, k, r: _nb(a[0], r[0], r[1], r[2]))(mdta_p1)
conv_to3 = _instrument("conv_to3", lambda a, k, r: _nb(a[0], a[1], r, k.get("residual")))(conv_to3)
conv_from3 = _instrument("conv_from3", lambda a, k, r: _nb(a[0], a[1], r, k.get("mask_y")))(conv_from3)
conv3_wgrad = _instrument("conv3_wgrad", lambda a, k, r: _nb(a[0], a[1]))(conv3_wgrad)
ln_stats = _instrument("ln_stats", lambda a, k, r: _nb(a[0], r))(ln_stats)
ln_fwd = _instrument("ln_fwd", lambda a, k, r: _nb(a[0], r[0]))(ln_fwd)
ln_bwd = _instrument("ln_bwd", lambda a, k, r: _nb(a[0], a[1], k.get("dy"), r))(ln_bwd)
dwconv = _instrument("dwconv", lambda a, k, r: _nb(a[0], r, k.get("dg"), k.get("g_out")))(dwconv)
dwconv_wgrad = _instrument("dwconv_wgrad", lambda a, k, r: _nb(a[0], a[1]))(dwconv_wgrad)
dwconv_bwd = _instrument("dwconv_bwd", lambda a, k, r: _nb(a[0], a[1], r))(dwconv_bwd)
gdfn_mid_bwd = _instrument("gdfn_mid_bwd", lambda a, k, r: _nb(a[0], a[1], r, k.get("g_out")))(gdfn_mid_bwd)
attn_fwd = _instrument("attn_fwd", lambda a, k, r: _nb(a[0], a[3]) * 2)(attn_fwd)
attn_bwd = _instrument("attn_bwd", lambda a, k, r: _nb(a[0], a[3]) * 3)(attn_bwd)
pixel_shuffle = _instrument("pixel_shuffle", lambda a, k, r: _nb(a[0], a[0]))(pixel_shuffle)
axpby = _instrument("axpby", lambda a, k, r: _nb(a[0], a[1] if len(a) > 1 else k.get("y"), r))(axpby)
channel_sum = _instrument("channel_sum", lambda a, k, r: _nb(a[0]))(channel_sum)
zero_ = _instrument("zero", lambda a, k, r: _nb(a[0]))(zero_)
linear_fwd = _instrument("linear_fwd", lambda a, k, r: _nb(a[0], a[1], r))(linear_fwd)
linear_dgrad = _instrument("linear_dgrad", lambda a, k, r: _nb(a[0], a[1], r))(linear_dgrad)
linear_wgrad = _instrument("linear_wgrad", lambda a, k, r: _nb(a[0], a[1], a[2]) + _nb(a[2]))(linear_wgrad)
cost_stage1 = _instrument("cost_stage1", lambda a, k, r: _nb(a[0], a[1], a[2], a[4]))(cost_stage1)
cost_stage2 = _instrument("cost_stage2", lambda a, k, r: _nb(a[0], a[1], a[2], a[3], a[4], a[6]))(cost_stage2)
sample_sumsq = _instrument("sample_sumsq", lambda a, k, r: _nb(a[0]))(sample_sumsq)
gp_coef = _instrument("gp_coef", lambda a, k, r: 0)(gp_coef)
signed_sum = _instrument("signed_sum", lambda a, k, r: 0)(signed_sum)
rmsprop = _instrument("rmsprop", lambda a, k, r: a[3] * 4 * 5)(rmsprop)
adam = _instrument("adam", lambda a, k, r: a[4] * 4 * 7)(adam)
rmsprop_h = _instrument("rmsprop", lambda a, k, r: a[3] * 4 * 5)(rmsprop_h)
adam_h = _instrument("adam", lambda a, k, r: a[4] * 4 * 7)(adam_h)
