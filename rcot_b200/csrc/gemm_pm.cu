// gemm_pm.cu -- "pixel-as-M" implicit GEMM on tcgen05:
//     out[b, n, p] = epilogue( sum_k A(b, p, k) * W[n, k] )
// M = 128 pixels of one image per CTA, N = output channels (up to 2 x 256 TMEM accumulators per
// CTA), K = Cin*ks*ks streamed through a shared-memory ring in chunks of 32.
//   * A operand: gathered by the CTA's threads (one thread = one pixel row) straight from the
//     NCHW fp32 activations -- optional concat of two tensors, optional LayerNorm prologue,
//     forward or transposed (data-gradient) conv geometry -- split to bf16 hi/lo and stored in the
//     no-swizzle K-major core-matrix layout.
//   * B operand: weights pre-packed in that layout (pack.cu), one cp.async.bulk per stage.
//   * D: fp32 in TMEM, read back with tcgen05.ld (thread = pixel, registers = channels) so the
//     epilogue (bias, LeakyReLU, sign mask, residual, accumulate) writes coalesced NCHW rows.
// Replaces the ATen conv2d calls listed in include/rcot_b200.h (rcot_pm_params).
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "tc.cuh"

namespace rcot {

constexpr int PM_MAX_STAGES = 4;

template <int KS, int MODE, int TERMS, bool LN>
__global__ void __launch_bounds__(128)
    pm_gemm_kernel(const rcot_pm_params p, const int Ktot, const int nk, const int BN, const int NSUB,
                   const int stages, const uint32_t tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[PM_MAX_STAGES], empty_bar[PM_MAX_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int TA = (TERMS > 1) ? 2 : 1;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z, pass = blockIdx.y;
  const int HWr = p.Hr * p.Wr, HWs = p.Hs * p.Ws;
  const int row_p = blockIdx.x * 128 + tid;
  const bool valid = row_p < HWr;
  const int ry = valid ? row_p / p.Wr : 0;
  const int rx = valid ? row_p - ry * p.Wr : 0;

  const uint32_t a_tile = 128 * KC * 2;
  const uint32_t b_tile = (uint32_t)BN * KC * 2;
  const uint32_t b_term = (uint32_t)NSUB * b_tile;
  const uint32_t stage_bytes = TA * a_tile + TA * b_term;

  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(128, BN);

  const float* src1 = p.in + (size_t)b * p.in_bs;
  const float* src2 = p.in2 ? p.in2 + (size_t)b * p.in2_bs : nullptr;
  float mu = 0.f, rstd = 0.f;
  if (LN && valid) {
    float2 st = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)b * HWs + row_p);
    mu = st.x;
    rstd = st.y;
  }
  const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)b * p.wpack_bs +
                        (size_t)pass * nk * (2 * b_term);

  for (int c = 0; c < nk; ++c) {
    const int s = c % stages, use = c / stages;
    if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
    uint8_t* st = smem + (size_t)s * stage_bytes;
    uint8_t* a_hi = st;
    uint8_t* a_lo = st + a_tile;
    uint8_t* b_st = st + TA * a_tile;
    if (tid == 0) {
      mbar_arrive_expect_tx(&full_bar[s], TA * b_term);
      bulk_g2s(b_st, wsrc + (size_t)c * (2 * b_term), TA * b_term, &full_bar[s]);
    }
#pragma unroll
    for (int k8 = 0; k8 < KC / 8; ++k8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = c * KC + k8 * 8 + i;
        float x = 0.f;
        if (valid && k < Ktot) {
          int ch, sy, sx;
          bool ok = true;
          if (KS == 1) {
            ch = k;
            sy = ry;
            sx = rx;
          } else {
            ch = k / (KS * KS);
            const int r = k - ch * (KS * KS);
            const int ky = r / KS, kx = r - ky * KS;
            if (MODE == 0) {
              sy = ry * p.stride + ky - p.pad;
              sx = rx * p.stride + kx - p.pad;
            } else {
              const int ty = ry + p.pad - ky, tx = rx + p.pad - kx;
              sy = ty / p.stride;
              sx = tx / p.stride;
              ok = (ty >= 0) && (tx >= 0) && (sy * p.stride == ty) && (sx * p.stride == tx);
            }
            ok = ok && ((unsigned)sy < (unsigned)p.Hs) && ((unsigned)sx < (unsigned)p.Ws);
          }
          if (ok) {
            const float* sp = (ch < p.C1) ? (src1 + (size_t)ch * HWs) : (src2 + (size_t)(ch - p.C1) * HWs);
            x = __ldg(sp + sy * p.Ws + sx);
            if (LN) x = (x - mu) * rstd * __ldg(p.ln_gamma + ch) + __ldg(p.ln_beta + ch);
          }
        }
        v[i] = x;
      }
      op_store8<TERMS>(a_hi, a_lo, tid, k8, v);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&full_bar[s], use & 1);
      tc_fence_after();
      for (int sub = 0; sub < NSUB; ++sub) {
        issue_stage<TERMS>(tmem + sub * BN, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_st) + sub * b_tile,
                           smem_u32(b_st) + b_term + sub * b_tile, idesc, c == 0);
      }
      tc_commit(&empty_bar[s]);
    }
  }
  if (tid == 0) tc_commit(&done_bar);
  mbar_wait(&done_bar, 0);
  tc_fence_after();

  // ---- epilogue: thread = pixel row, registers = 16 output channels at a time
  const uint32_t lane_base = tmem_lane_base(tmem);
  float* outb = p.out + (size_t)b * p.out_bs + (size_t)p.out_coff * HWr;
  const float* maskb = p.mask_y ? p.mask_y + (size_t)b * p.mask_bs : nullptr;
  const float* resb = p.residual ? p.residual + (size_t)b * p.res_bs : nullptr;
  for (int sub = 0; sub < NSUB; ++sub) {
    const int nbase = (pass * NSUB + sub) * BN;
    if (nbase >= p.N) break;
    for (int n0 = 0; n0 < BN; n0 += 16) {
      if (nbase + n0 >= p.N) break;
      float v[16];
      tmem_ld16(lane_base + sub * BN + n0, v);
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = nbase + n0 + i;
          if (n < p.N) {
            float y = v[i];
            if (p.bias) y += __ldg(p.bias + n);
            if (p.act) y = y > 0.f ? y : y * p.slope;
            const size_t idx = (size_t)n * HWr + row_p;
            if (maskb) y *= (__ldg(maskb + idx) > 0.f) ? 1.f : p.slope;
            if (resb) y += __ldg(resb + idx);
            if (p.accumulate) y += outb[idx];
            outb[idx] = y;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <int KS, int MODE, int TERMS, bool LN>
static int launch_pm(const rcot_pm_params& p, cudaStream_t stream) {
  const int K = (p.C1 + p.C2) * KS * KS;
  const int nk = cdiv(K, KC);
  NPlan pl = make_nplan(p.N);
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  const size_t stage_bytes = (size_t)TA * (128 * KC * 2) + (size_t)TA * pl.NSUB * pl.BN * KC * 2;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > PM_MAX_STAGES) stages = PM_MAX_STAGES;
  if (stages > nk) stages = nk < 1 ? 1 : nk;
  if (stages < 1) stages = 1;
  const size_t smem = stages * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pm_gemm_kernel<KS, MODE, TERMS, LN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("pm_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(cdiv((long)p.Hr * p.Wr, 128), pl.passes, p.B);
  pm_gemm_kernel<KS, MODE, TERMS, LN><<<grid, 128, smem, stream>>>(p, K, nk, pl.BN, pl.NSUB, stages,
                                                                     tmem_cols_pow2(pl.NSUB * pl.BN));
  return check_launch("pm_gemm");
}

}  // namespace rcot

extern "C" int rcot_pm_gemm(const rcot_pm_params* pp, rcot_stream_t stream_) {
  using namespace rcot;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RCOT_REQUIRE(pp != nullptr, "pm_gemm: null params");
  const rcot_pm_params& p = *pp;
  RCOT_REQUIRE(p.in && p.out && p.wpack, "pm_gemm: null tensor pointer");
  RCOT_REQUIRE(p.B > 0 && p.C1 > 0 && p.C2 >= 0 && p.N > 0, "pm_gemm: bad sizes B=%d C1=%d C2=%d N=%d", p.B, p.C1,
               p.C2, p.N);
  RCOT_REQUIRE((p.C2 == 0) == (p.in2 == nullptr), "pm_gemm: in2/C2 mismatch");
  RCOT_REQUIRE(p.Hs > 0 && p.Ws > 0 && p.Hr > 0 && p.Wr > 0, "pm_gemm: bad spatial sizes");
  RCOT_REQUIRE(p.terms == 1 || p.terms == 3, "pm_gemm: terms must be 1 or 3");
  RCOT_REQUIRE(p.mode == 0 || p.mode == 1, "pm_gemm: mode must be 0 or 1");
  RCOT_REQUIRE(p.B <= 65535, "pm_gemm: batch too large for grid.z");
  const bool ln = p.ln_stats != nullptr;
  if (ln) RCOT_REQUIRE(p.ks == 1 && p.ln_gamma && p.ln_beta, "pm_gemm: LayerNorm prologue needs ks==1, gamma, beta");
  if (p.ks == 1)
    RCOT_REQUIRE(p.stride == 1 && p.pad == 0 && p.Hs == p.Hr && p.Ws == p.Wr, "pm_gemm: 1x1 needs stride 1, pad 0");
  RCOT_REQUIRE(p.stride >= 1, "pm_gemm: stride must be >= 1");
#define PM_DISPATCH(KS, MODE, LN)                                           \
  return (p.terms == 3) ? launch_pm<KS, MODE, 3, LN>(p, stream) : launch_pm<KS, MODE, 1, LN>(p, stream)
  switch (p.ks) {
    case 1:
      if (ln) { PM_DISPATCH(1, 0, true); }
      PM_DISPATCH(1, 0, false);
    case 3:
      if (p.mode == 0) { PM_DISPATCH(3, 0, false); }
      PM_DISPATCH(3, 1, false);
    case 4:
      if (p.mode == 0) { PM_DISPATCH(4, 0, false); }
      PM_DISPATCH(4, 1, false);
    case 5:
      if (p.mode == 0) { PM_DISPATCH(5, 0, false); }
      PM_DISPATCH(5, 1, false);
    default:
      set_error("pm_gemm: unsupported kernel size %d", p.ks);
      return RCOT_ERR_ARG;
  }
#undef PM_DISPATCH
}
