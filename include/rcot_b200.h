/* rcot_b200.h -- C ABI of librcot_b200.so (sm_100a only; no CPU / other-arch fallback).
 *
 * The reference (xl-tang3/RCOT) has no FFI: its hot path is PyTorch-eager Python.  Each entry
 * point below replaces the ATen call sequence of the reference lines it cites; the Python host
 * (rcot_b200/ops.py) binds them with ctypes and INTEGRATION.md shows the binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *  - all tensors are fp32, NCHW, contiguous per image; `*_bs` is the batch stride in elements,
 *    so channel slices / concatenated views need no copies
 *  - the caller owns every buffer (PyTorch caching allocator); kernels never allocate
 *  - every call is asynchronous on `stream` and CUDA-graph capturable
 *  - return 0 on success, <0 on error; rcot_last_error() gives the message (thread-local)
 *  - `terms`: 3 = bf16x3 split products on tcgen05 (fp32-class accuracy, the parity mode),
 *             1 = single bf16 product (bf16 compute)
 */
#ifndef RCOT_B200_H
#define RCOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* rcot_stream_t;

int rcot_version(void);
const char* rcot_last_error(void);
int rcot_check_device(void);

/* tcgen05 bring-up: D[128 x N] = A[128 x K] * B[N x K]^T */
int rcot_selftest_tc(const float* A, const float* B, float* D, int N, int K, int terms, rcot_stream_t stream);
/* same product with the A operand resident in TENSOR MEMORY (tcgen05.mma "TS" form; single bf16 term): bring-up of the
 * TMEM operand layout for fused kernels that keep a tile's A operand on chip across many small MMAs */
int rcot_selftest_tmem_a(const float* A, const float* B, float* D, int N, int K, int variant, rcot_stream_t stream);

/* ---------------------------------------------------------------- packed weights
 * B-operand image of a [N x K] matrix in the tcgen05 no-swizzle K-major layout, bf16 hi+lo,
 * tiled [pass][k-chunk][term][sub-tile][BN x 32].  Element (n, k) is read from
 *   src[(k / R) * s_kouter + (k % R) + n * s_n]
 * which covers conv forward (R > K, s_n = K), conv dgrad (R = ks*ks, s_kouter = Cin*ks*ks,
 * s_n = ks*ks) and transposed 1x1 (R = 1, s_kouter = Cin, s_n = 1). */
typedef struct {
  const float* src;
  void* dst;
  int32_t N, K;
  int32_t R, s_kouter, s_n;
  int32_t s_kinner;     /* stride of (k % R); 0 means 1 */
} rcot_pack_desc;

size_t rcot_packed_bytes(int N, int K);
/* descs: DEVICE array of n descriptors; one launch packs them all. max_elems = max over descs of
 * padded N * padded K (use rcot_packed_bytes / 4). */
int rcot_pack_weights(const rcot_pack_desc* descs, int n, size_t max_elems, rcot_stream_t stream);

/* ---------------------------------------------------------------- pixel-as-M GEMM
 * out[b, coff + n, p] = epilogue( sum_k A(b, p, k) * W[n, k] ), rows p = pixels of one image.
 * Replaces F.conv2d for: 1x1 convs incl. the LayerNorm in front of them
 * (Net_Restormer.py:25,27,73,78,186-189,282-316), dense 3x3 glue convs (:90,107,117,326), the
 * F_net conv stack (:443-489) and, with mode=1, their data gradients. */
typedef struct {
  const float* in;      /* first C1 channels of the gather source */
  const float* in2;     /* next C2 channels (concat without a copy), may be NULL */
  int64_t in_bs, in2_bs;
  int32_t C1, C2;
  int32_t Hs, Ws;       /* source spatial size */
  int32_t Hr, Wr;       /* row-space spatial size (output pixels; input pixels for mode=1) */
  int32_t B;
  int32_t ks, stride, pad;
  int32_t mode;         /* 0 forward gather, 1 transposed (dgrad) gather */
  const float* ln_stats;  /* [B, Hs*Ws, 2] (mean, rstd) -> LayerNorm prologue, ks==1 only */
  const float* ln_gamma;
  const float* ln_beta;
  const void* wpack;
  int64_t wpack_bs;     /* bytes between per-image weights, 0 = shared */
  int32_t N;
  int32_t terms;
  float* out;
  int64_t out_bs;
  int32_t out_coff;
  int32_t act;          /* 1: LeakyReLU(slope) after bias */
  float slope;
  int32_t accumulate;   /* out += result */
  const float* bias;    /* [N] or NULL */
  const float* mask_y;  /* same indexing as out: result *= (mask_y > 0 ? 1 : slope) */
  int64_t mask_bs;
  const float* residual; /* same indexing as out (without coff): result += residual */
  int64_t res_bs;
  float* stats_out;     /* optional [B, Hr*Wr, 2]: LayerNorm (mean, rstd) over the N output channels of every pixel,
                           computed in the epilogue (needs N <= 256, i.e. one pass) for the LN that consumes `out` */
  int32_t debug;        /* A/B knobs of the TMA-staged 1x1 variant (0 = all on): bit 0 loads the LayerNorm statistics at
                           the tile's start instead of one tile ahead, bit 1 disables the residual L2 prefetch */
  int32_t tap_major;    /* ks > 1: K is ordered (ky, kx, channel); needs C1 % 16 == 0 and no concat: a producer
                           thread's 16 K indices are 16 channels at ONE tap: a strided read like the 1x1 case */
  int32_t in_bf16;      /* bf16-storage mode: `in` is a bf16 tensor (1x1, no LayerNorm, no concat, H*W % 128 == 0): the A
                           operand is exact in bf16, so its lo term and the fp32->bf16 split disappear */
  int32_t out_bf16;     /* `out` is a bf16 tensor (1x1, plain epilogue: no bias/activation/mask/residual/accumulate/stats) */
  /* LayerNorm-BACKWARD epilogue (1x1, N = C <= 256; Net_Restormer.py:186-189 differentiated): the GEMM result is
   * dz = dL/dLN(x); out = [residual +] LN'(dz) with x / (mean, rstd) / gamma of that LayerNorm, and
   * lnb_dgamma += sum dz * xhat, lnb_dbeta += sum dz.  lnb_x == NULL: off. */
  const float* lnb_x;
  int64_t lnb_x_bs;
  const float* lnb_stats;
  const float* lnb_gamma;
  float* lnb_dgamma;
  float* lnb_dbeta;
} rcot_pm_params;

int rcot_pm_gemm(const rcot_pm_params* p, rcot_stream_t stream);

/* ---------------------------------------------------------------- pixel-as-K GEMM
 * out[(b,) m, n] += sum over pixels q of A[b, m, q] * Bg(b, n, q)
 * Weight gradients of every conv (dW = dOut * im2col(In)^T), and MDTA's per-image Gram
 * q k^T / dy v^T (Net_Restormer.py:42 and its backward). */
typedef struct {
  const float* a;       /* [B, CA, Ha*Wa] (bf16 when a_bf16) */
  int64_t a_bs;
  int32_t CA;
  const float* b;       /* [B, CB, Hb, Wb] gather source */
  const float* b2;      /* concat continuation, may be NULL */
  int64_t b_bs, b2_bs;
  int32_t CB1, CB2;
  int32_t Ha, Wa, Hb, Wb;
  int32_t B;
  int32_t ks, stride, pad;
  const float* ln_stats; /* LayerNorm applied to b on the fly (ks==1) */
  const float* ln_gamma;
  const float* ln_beta;
  int32_t per_image;    /* 1: out is [B, CA, N] (Gram); 0: reduce over the batch */
  int32_t terms;
  float* out;           /* accumulated with atomics: caller zeroes or accumulates into .grad */
  int64_t out_bs;       /* per-image stride of out (per_image=1) */
  int32_t ldo;
  int32_t groups;       /* >1: a channels g*CA.., b channels g*CB1.., out + g*out_gs (MDTA heads) */
  int64_t out_gs;
  int32_t a_bf16, b_bf16; /* bf16-storage mode: a / b is a bf16 tensor (1x1, aligned shapes with H*W % 32 == 0; b only
                             without LayerNorm): the operand is exact in bf16, its lo term disappears */
} rcot_pk_params;

int rcot_pk_gemm(const rcot_pk_params* p, rcot_stream_t stream);

/* ---------------------------------------------------------------- device-side training patches (data path)
 * util/dataset_utils.py:215-278 (TrainDataset.__getitem__), util/image_utils.py:59-65,133-182,
 * util/degradation_utils.py:21-27 for a whole batch: centre crop to multiples of 16, P x P crop at (y0, x0),
 * augmentation `mode` (0..7), then either Gaussian noise on the uint8 grid (sigma > 0; `noise` = [B,P,P,3] float32
 * standard normals) or the paired degraded image (sigma == 0).  Images: uint8 HWC in one pool buffer.
 * Outputs: degraded, clean = fp32 NCHW [B,3,P,P] in [0,1]. */
typedef struct {
  int64_t clean_off, deg_off;   /* byte offsets of the clean / degraded image inside the pool (deg unused if sigma>0) */
  int32_t H, W;                 /* image size before the centre crop                                                */
  int32_t y0, x0;               /* random-crop origin inside the centre-cropped image                               */
  int32_t mode;                 /* augmentation 0..7 (util/image_utils.py:133-163)                                   */
  float sigma;                  /* 15 / 25 / 50 for the denoise tasks, 0 for paired tasks                            */
} rcot_patch_desc;
int rcot_make_patches(const uint8_t* pool, const rcot_patch_desc* desc, const float* noise, float* degraded,
                      float* clean, int B, int P, rcot_stream_t stream);

/* ---------------------------------------------------------------- fused GDFN forward (one kernel)
 * Net_Restormer.py:80-85 (+ the block's norm2 and residual, :212-213):
 *     y = [x +] project_out( gelu(dwconv(project_in(LN(x)))[:hid]) * dwconv(project_in(LN(x)))[hid:] )
 * with the hidden tensor kept on chip (8x16-pixel tiles with a 1-pixel halo, hidden dimension walked in slices of 16
 * channel pairs, both 1x1 convs on tcgen05).  Built for C in {48, 96} (the levels that carry 84 % of the bytes), H % 8
 * == 0, W % 16 == 0; other shapes use the three-launch path.  Weights come as one blob made by rcot_gdfn_pack from
 * project_in.weight [2*hid, C], dwconv.weight [2*hid, 9] and project_out.weight [C, hid] (re-pack after each step).
 * Optional save_u [B, 2*hid, H, W] / save_g [B, hid, H, W] receive the hidden tensors for a backward that wants them. */
typedef struct {
  const float* x;            /* [B, C, H, W], per-image block contiguous                     */
  int64_t x_bs;
  const float* ln_stats;     /* [B, H*W, 2] (mean, rstd) or NULL = no LayerNorm              */
  const float* ln_gamma;
  const float* ln_beta;
  const void* wblob;         /* rcot_gdfn_pack output                                        */
  float* y;                  /* [B, C, H, W]                                                 */
  int64_t y_bs;
  float* stats_out;          /* optional: LayerNorm statistics of y                          */
  float* save_u;             /* optional                                                      */
  int64_t u_bs;
  float* save_g;             /* optional (16-byte aligned)                                    */
  int64_t g_bs;
  int32_t B, C, H, W, hid;
  int32_t residual;          /* y = x + out                                                   */
  int32_t debug;             /* measurement knobs (0 in production): 1 skip MMAs, 2 skip stencil, 4 skip drain stores, 16 cycle counters */
} rcot_gdfn_params;
int rcot_gdfn_supported(int C, int H, int W);
size_t rcot_gdfn_blob_bytes(int C, int hid);
int rcot_gdfn_pack(const float* w_in, const float* w_dw, const float* w_out, void* blob, int C, int hid,
                   rcot_stream_t stream);
int rcot_gdfn_fwd(const rcot_gdfn_params* p, rcot_stream_t stream);
/* Measurement aid (not part of the reference surface): per-warp wait / section cycle counters of CTA 0 of the last
 * rcot_gdfn_fwd launched with debug & 16, as [22 warps][8] uint64 (n = 176).  Synchronises the device. */
int rcot_gdfn_profile_read(unsigned long long* out, int n);

/* ---------------------------------------------------------------- convolutions that END in three channels
 * Direct FP32 kernel (csrc/conv3.cu) for Net_Restormer.py:326 (output conv 96 -> 3, forward: weight [3, Cin, k, k]) and
 * for the data gradient of a conv that STARTS from three channels (:117 patch_embed, :443 F_net features.0; dgrad = 1:
 * `in` is dL/dy with Cin = the conv's Cout channels, weight [Cin, 3, k, k], taps flipped).  Stride 1, pad (k-1)/2,
 * k in {3, 5}; optional residual [B, 3, H, W] added to the result (the `+ inp_img` of T_net.forward). */
int rcot_conv_to3(const float* in, int64_t in_bs, const float* weight, int dgrad, float* out, int64_t out_bs,
                  const float* residual, int64_t res_bs, int B, int Cin, int H, int W, int ks, rcot_stream_t stream);

/* Direct FP32 kernel for a conv that STARTS from three channels (Net_Restormer.py:117 patch_embed 3 -> 48 k3, :443 F_net
 * features.0 3 -> 64 k5): weight [Cout, 3, k, k], stride 1, pad (k-1)/2, k in {3, 5}; optional bias [Cout] and
 * LeakyReLU(slope) (act = 1), or -- the bias-free tangent pass of the gradient penalty -- the LeakyReLU-derivative mask
 * of a previous forward's output mask_y [B, Cout, H, W]: out *= (mask_y > 0 ? 1 : slope). */
int rcot_conv_from3(const float* in, int64_t in_bs, const float* weight, const float* bias, float* out, int64_t out_bs,
                    const float* mask_y, int64_t mask_bs, int act, float slope, int B, int Cout, int H, int W, int ks,
                    rcot_stream_t stream);

/* Weight gradient of a stride-1 'same' conv with three channels on one side (k in {3, 5}), ACCUMULATED into dw:
 * from3 = 0: the conv ends in 3 channels (weight / dw [3, Cm, k, k]): many = its input [B, Cm, H, W], three = dL/dy;
 * from3 = 1: the conv starts from 3 channels (weight / dw [Cm, 3, k, k]): many = dL/dy [B, Cm, H, W], three = its input. */
int rcot_conv3_wgrad(const float* many, int64_t many_bs, const float* three, int64_t three_bs, float* dw, int from3, int B,
                     int Cm, int H, int W, int ks, rcot_stream_t stream);

/* ---------------------------------------------------------------- MDTA phase 1 as ONE kernel (csrc/mdta_fused.cu)
 * Net_Restormer.py:29-41 (qkv 1x1 conv of LN(x), depthwise 3x3, q k^T and the row norms of F.normalize) with pre, q
 * and k kept on chip: per 8x16-pixel tile (+1-pixel halo) the 3C channels are walked in slices of 32 (tcgen05 GEMM from
 * a TMEM-resident LN(x) operand -> shared-memory stencil), v goes to HBM, q / k become bf16 hi/lo operand rows of a
 * per-tile Gram MMA that accumulates in TMEM across the tiles of an image.  Built for C in {48, 96}, H % 8 == 0,
 * W % 16 == 0.  G [B, heads, c, c] and sumsq [B, 2C] are ACCUMULATED (zero them first); they and v are exactly what
 * rcot_attn_fwd and the y = x + M v GEMM consume.  Weights: one blob made by rcot_mdta_p1_pack from qkv.weight [3C, C]
 * and qkv_dwconv.weight [3C, 9].  Optional save_pre [B, 3C, H, W] / save_qk [B, 2C, H, W] feed the unfused backward. */
typedef struct {
  const float* x;            /* [B, C, H, W], per-image block contiguous                     */
  int64_t x_bs;
  const float* ln_stats;     /* [B, H*W, 2] (mean, rstd) or NULL = no LayerNorm              */
  const float* ln_gamma;
  const float* ln_beta;
  const void* wblob;         /* rcot_mdta_p1_pack output                                     */
  float* v;                  /* [B, C, H, W] (16-byte aligned; may be the v part of a qkv tensor) */
  int64_t v_bs;
  float* G;                  /* [B, heads, c, c]  +=                                         */
  float* sumsq;              /* [B, 2C]           +=  (q channels then k channels)           */
  float* save_pre;           /* optional                                                     */
  int64_t pre_bs;
  float* save_qk;            /* optional (may be the q, k part of a qkv tensor)              */
  int64_t qk_bs;
  int32_t B, C, H, W, heads;
  int32_t debug;             /* measurement knobs (0 in production): 1 skip MMAs, 2 skip stencil, 4 skip drain stores */
} rcot_mdta_p1_params;
int rcot_mdta_p1_supported(int C, int H, int W, int heads);
size_t rcot_mdta_p1_blob_bytes(int C);
int rcot_mdta_p1_pack(const float* w_qkv, const float* w_dw, void* blob, int C, rcot_stream_t stream);
int rcot_mdta_p1(const rcot_mdta_p1_params* p, rcot_stream_t stream);

/* ---------------------------------------------------------------- LayerNorm over channels
 * Net_Restormer.py:173-200 (WithBias_LayerNorm on the 'b (h w) c' view): stats[b, p] = (mean, rstd)
 * with biased variance and eps 1e-5; the normalisation itself is applied as a GEMM prologue. */
int rcot_ln_stats(const float* x, int64_t x_bs, int B, int C, int HW, float* stats, rcot_stream_t stream);
/* Stand-alone LayerNorm forward (Net_Restormer.py:186-189,198-200): y = (x-mu)*rstd*gamma+beta per pixel over C;
 * also writes the (mu, rstd) pairs rcot_ln_bwd needs. */
int rcot_ln_fwd(const float* x, int64_t x_bs, const float* gamma, const float* beta, float* y, int64_t y_bs, int B, int C,
                int HW, float* stats, rcot_stream_t stream);
/* dx = [dy +] LN'(dz); dgamma += sum dz*xhat; dbeta += sum dz  (dy may be NULL; dx may alias dy or dz) */
int rcot_ln_bwd(const float* dz, int64_t dz_bs, const float* x, int64_t x_bs, const float* stats,
                const float* gamma, const float* dy, int64_t dy_bs, float* dx, int64_t dx_bs, float* dgamma,
                float* dbeta, int B, int C, int HW, rcot_stream_t stream);

/* ---------------------------------------------------------------- depthwise 3x3 (pad 1, no bias)
 * Net_Restormer.py:26 (qkv_dwconv), :75 + :81-83 (GDFN dwconv, chunk, gelu(x1)*x2).
 * mode 0: out[ch] = dw(in[ch]); flip=1 gives the data gradient; sumsq[b*nsq+ch] += sum_p out^2 (ch<nsq)
 * mode 1: out[j] = gelu_erf(dw(in[j])) * dw(in[j+hid])                      (Cn = 2*hid, out has hid)
 * mode 2: gate backward: out[j] = dg*b*gelu'(a), out[j+hid] = dg*gelu(a); g_out[j] = gelu(a)*b (optional) */
typedef struct {
  const float* in;
  int64_t in_bs;
  const float* w;       /* [Cn, 1, 3, 3] */
  float* out;
  int64_t out_bs;
  int32_t B, Cn, H, W;
  int32_t mode, flip, hid, nsq;
  const float* dg;
  int64_t dg_bs;
  float* g_out;
  int64_t g_bs;
  float* sumsq;
  int32_t bf16;         /* 1: in / out / dg / g_out are bf16 tensors (bf16-storage mode of the hidden tensors; fp32
                           arithmetic).  Needs the aligned geometry of training patches (W % 4 == 0, even H). */
} rcot_dw_params;
int rcot_dwconv3x3(const rcot_dw_params* p, rcot_stream_t stream);
/* dw[ch, k] += sum_{b,p} dout[b,ch,p] * in[b,ch,p+off_k] */
int rcot_dwconv3x3_wgrad(const float* in, int64_t in_bs, const float* dout, int64_t dout_bs, float* dw, int B,
                         int Cn, int H, int W, rcot_stream_t stream);

/* both halves of the depthwise backward in one pass: din = dw^T(dout), dw += corr(in, dout) */
int rcot_dwconv3x3_bwd(const float* in, int64_t in_bs, const float* dout, int64_t dout_bs, const float* w, float* din,
                       int64_t din_bs, float* dw, int B, int Cn, int H, int W, rcot_stream_t stream);

/* same, with in / dout / din stored as bf16 when bf16 != 0 (dw stays fp32) */
int rcot_dwconv3x3_bwd_t(const void* in, int64_t in_bs, const void* dout, int64_t dout_bs, const float* w, void* din,
                         int64_t din_bs, float* dw, int B, int Cn, int H, int W, int bf16, rcot_stream_t stream);

/* GDFN middle backward in one pass (Net_Restormer.py:81-83 backward): with a = dw(u[j]),
 * b = dw(u[j+hid]): da = dg*b*gelu'(a), db = dg*gelu(a); du = dw^T([da; db]); dw += corr(u, [da; db]);
 * g_out[j] = gelu(a)*b (optional, may be NULL).  Same arithmetic as mode 2 followed by rcot_dwconv3x3_bwd, without
 * the [da; db] round trip through HBM.  Requires W % 32 == 0, H % 4 == 0 and 16-byte aligned u / dg / du / g_out. */
int rcot_gdfn_mid_bwd(const float* u, int64_t u_bs, const float* dg, int64_t dg_bs, const float* w, float* du,
                      int64_t du_bs, float* dw, float* g_out, int64_t g_bs, int B, int hid, int H, int W,
                      rcot_stream_t stream);

/* ---------------------------------------------------------------- MDTA small-matrix steps
 * Net_Restormer.py:39-49: normalize(q), normalize(k), softmax(q k^T * temperature), attn @ v,
 * project_out -- folded into the per-image matrix M = W_out * blockdiag(A). */
typedef struct {
  int32_t B, C, heads, reserved;
  const float* G;            /* [B, heads, c, c] raw q k^T                       (fwd in)  */
  const float* sumsq;        /* [B, 2C] sum over pixels of q^2 then k^2                    */
  const float* temperature;  /* [heads]                                                    */
  const float* w_out;        /* [C, C] project_out.weight                                  */
  float* A;                  /* [B, heads, c, c] softmax probabilities   (fwd out, bwd in) */
  float* Gt;                 /* [B, heads, c, c] normalised Gram         (fwd out, bwd in) */
  void* Mpack;               /* per-image packed M   (N=C, K=C)           (fwd out)        */
  void* MTpack;              /* per-image packed M^T, may be NULL         (fwd out)        */
  int64_t pack_bs;           /* bytes between images = rcot_packed_bytes(C, C)             */
  const float* P;            /* [B, C, C] dy v^T                          (bwd in)         */
  float* dw_out;             /* [C, C] +=                                  (bwd out)       */
  float* dtemperature;       /* [heads] +=                                 (bwd out)       */
  void* W12pack;             /* per-image packed [2C x 2C], zero-initialised by the caller */
  int64_t pack12_bs;         /* = rcot_packed_bytes(2C, 2C)                                */
  float* dA;                 /* [B, heads, c, c] scratch, ZEROED by the caller before rcot_attn_bwd (its first
                                kernel accumulates W_out^T P there in row chunks, its second consumes it) */
} rcot_attn_params;
int rcot_attn_fwd(const rcot_attn_params* p, rcot_stream_t stream);
int rcot_attn_bwd(const rcot_attn_params* p, rcot_stream_t stream);

/* ---------------------------------------------------------------- data movement / small reductions */
/* PixelShuffle(2) (inverse=0) / PixelUnshuffle(2) (inverse=1), Net_Restormer.py:91,108.
 * C,H,W describe the [4C,H,W] side; the other side is [C,2H,2W]. */
int rcot_pixel_shuffle(const float* in, int64_t in_bs, float* out, int64_t out_bs, int B, int C, int H, int W,
                       int inverse, rcot_stream_t stream);
/* out = a*x + b*y (y may be NULL); with a_vec: out = a_vec[b]*x + (1-a_vec[b])*y (trainer.py:286) */
int rcot_axpby(float* out, int64_t out_bs, const float* x, int64_t x_bs, const float* y, int64_t y_bs, float a,
               float b, const float* a_vec, int B, int64_t n, rcot_stream_t stream);
/* out[c] += sum_{b,p} x[b,c,p] (bias gradients) */
int rcot_channel_sum(const float* x, int64_t x_bs, float* out, int B, int C, int HW, rcot_stream_t stream);
/* cudaMemsetAsync(ptr, 0, bytes) on the stream (graph-capturable) */
int rcot_zero(void* ptr, size_t bytes, rcot_stream_t stream);

/* ---------------------------------------------------------------- F_net fully connected tail
 * Net_Restormer.py:496-498,512-520 (fc, fc1, LeakyReLU, fc2) -- fp32 CUDA-core, row-major [B,K] x [O,K]^T.
 * mask (same shape as the result) multiplies by (mask > 0 ? 1 : slope): LeakyReLU derivative taken
 * from the saved post-activation. */
int rcot_linear_fwd(const float* x, const float* W, const float* bias, const float* mask, float* y, int B, int K,
                    int O, int act, float slope, rcot_stream_t stream);
int rcot_linear_dgrad(const float* dy, const float* W, const float* mask, float* dx, int B, int K, int O, float slope,
                      rcot_stream_t stream);
/* dW += dy^T x ; db += sum_b dy (db may be NULL) */
int rcot_linear_wgrad(const float* dy, const float* x, float* dW, float* db, int B, int K, int O,
                      rcot_stream_t stream);

/* ---------------------------------------------------------------- transport cost (trainer.py:320-343)
 * stage 1, per (image, channel): acc[0] += sum res^2, acc[1] += Fourier penalty (sum over the batch),
 * acc[2] += sum |out - target| (target may be NULL = unpaired); gfou = d(Fourier)/d(res).
 * de_id: int64 [B] on the device (de_id < 3 selects the mean|F|^2/2 branch).  P: power of two <= 128.
 * stage 2: dout = dF - sigma*(res/(N*rmse) + gfou) + Sigma*sign(out-target)/N, rmse = sqrt(acc[0]/N),
 * N = n_global (elements of the GLOBAL batch; acc[0] must be all-reduced first when data parallel). */
int rcot_cost_stage1(const float* out, const float* degraded, const float* target, const int64_t* de_id, float* gfou,
                     float* acc, int B, int P, rcot_stream_t stream);
int rcot_cost_stage2(const float* out, const float* degraded, const float* target, const float* gfou, const float* dF,
                     const float* acc, float* dout, float sigma, float Sigma, double n_global, int64_t n,
                     rcot_stream_t stream);

/* ---------------------------------------------------------------- gradient penalty helpers (trainer.py:300-305) */
int rcot_sample_sumsq(const float* x, float* out, int B, int64_t n, rcot_stream_t stream); /* out[b] += sum x[b,:]^2 */
/* coef[b] = 20/B_global * (|g_b|-1)/|g_b| ; loss[0] += 10/B_global * sum_b (|g_b|-1)^2 */
int rcot_gp_coef(const float* sumsq, float* coef, float* loss, int B, int B_global, rcot_stream_t stream);
/* out[0] += scale * (sum_{i>=n_neg} x[i] - sum_{i<n_neg} x[i]) */
int rcot_signed_sum(const float* x, float* out, int n, int n_neg, float scale, rcot_stream_t stream);

/* ---------------------------------------------------------------- optimizers on flat buffers
 * torch.optim.RMSprop / Adam defaults as used at trainer.py:121-126; g is multiplied by gscale first. */
int rcot_rmsprop(float* p, const float* g, float* sq, int64_t n, float lr, float alpha, float eps, float gscale,
                 rcot_stream_t stream);
int rcot_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
              int step, float gscale, rcot_stream_t stream);
/* same, with the step-dependent scalars in DEVICE memory so a captured CUDA graph can be replayed while the
 * schedule moves: hyper = {lr, 1-b1^t, 1-b2^t} (RMSprop reads hyper[0] only); effective lr = hyper[0]*lr_mult */
int rcot_rmsprop_h(float* p, const float* g, float* sq, int64_t n, const float* hyper, float lr_mult, float alpha,
                   float eps, float gscale, rcot_stream_t stream);
int rcot_adam_h(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float lr_mult, float b1,
                float b2, float eps, float gscale, rcot_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RCOT_B200_H */
