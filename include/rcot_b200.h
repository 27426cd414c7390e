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
  int32_t reserved;
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
} rcot_pm_params;

int rcot_pm_gemm(const rcot_pm_params* p, rcot_stream_t stream);

/* ---------------------------------------------------------------- pixel-as-K GEMM
 * out[(b,) m, n] += sum over pixels q of A[b, m, q] * Bg(b, n, q)
 * Weight gradients of every conv (dW = dOut * im2col(In)^T), and MDTA's per-image Gram
 * q k^T / dy v^T (Net_Restormer.py:42 and its backward). */
typedef struct {
  const float* a;       /* [B, CA, Ha*Wa] */
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
  int64_t out_bs;
  int32_t ldo;
  int32_t reserved;
} rcot_pk_params;

int rcot_pk_gemm(const rcot_pk_params* p, rcot_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RCOT_B200_H */
