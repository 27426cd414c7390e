// gemm_pk.cu -- "pixel-as-K" GEMM on tcgen05:
//     out[(b,g,) m, n] += sum over pixels q of  a[b, g*CA + m, q] * Bg(b, g, n, q)
// with n = (cb, ky, kx) and Bg = b[b, g*CB + cb, qy*stride + ky - pad, qx*stride + kx - pad] (zero outside).
// Both operands are K-major in NCHW as they lie (pixels are contiguous), so threads only convert
// fp32 -> bf16 hi/lo into the core-matrix layout; 128 x BN fp32 accumulators live in TMEM; split-K
// CTAs add their partial tiles with fp32 atomics.
// Covers every weight gradient (dW = dOut * im2col(In)^T; the 1x1 ones with the LayerNorm applied to
// In on the fly) and MDTA's per-image channel Grams q k^T and dy v^T (Net_Restormer.py:42 and its
// backward).
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "tc.cuh"

namespace rcot {

constexpr int PK_THREADS = 256;
constexpr int PK_STAGES = 2;

template <int TERMS, bool GENERAL, bool LN>
__global__ void __launch_bounds__(PK_THREADS)
    pk_gemm_kernel(const rcot_pk_params p, const int BN, const int nt, const int cpi, const int per_cta,
                   const int total_chunks, const uint32_t tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t empty_bar[PK_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int mt_i = blockIdx.x / nt, nt_i = blockIdx.x - mt_i * nt;
  const int m0 = mt_i * 128, n0 = nt_i * BN;
  const int g = blockIdx.z % p.groups;
  const int bz = blockIdx.z / p.groups;  // image index when per_image, else 0
  const int HWa = p.Ha * p.Wa, HWb = p.Hb * p.Wb;
  const int KK = p.ks * p.ks;
  const int Ntot = (p.CB1 + p.CB2) * KK;

  const uint32_t a_tile = 128 * KC * 2, b_tile = (uint32_t)BN * KC * 2;
  const uint32_t stage_bytes = TA * (a_tile + b_tile);

  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < PK_STAGES; ++s) mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(128, BN);

  const int c_begin = blockIdx.y * per_cta;
  int c_end = c_begin + per_cta;
  if (c_end > total_chunks) c_end = total_chunks;

  for (int gc = c_begin; gc < c_end; ++gc) {
    const int it = gc - c_begin;
    const int s = it % PK_STAGES, use = it / PK_STAGES;
    if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
    int b, q0;
    if (p.per_image) {
      b = bz;
      q0 = gc * KC;
    } else {
      b = gc / cpi;
      q0 = (gc - b * cpi) * KC;
    }
    uint8_t* st = smem + (size_t)s * stage_bytes;
    uint8_t* a_hi = st;
    uint8_t* a_lo = st + a_tile;
    uint8_t* b_hi = st + TA * a_tile;
    uint8_t* b_lo = b_hi + b_tile;

    // ---- A operand: rows = channels m0.., k = pixels q0..q0+31
    const float* ab = p.a + (size_t)b * p.a_bs + (size_t)g * p.CA * HWa;
    for (int task = tid; task < 128 * (KC / 8); task += PK_THREADS) {
      const int r = task >> 2, k8 = task & 3;
      const int m = m0 + r, q = q0 + k8 * 8;
      float v[8];
      if (m < p.CA && q + 8 <= HWa && !GENERAL) {
        const float4* src = reinterpret_cast<const float4*>(ab + (size_t)m * HWa + q);
        const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
        v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
        v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (m < p.CA && q + i < HWa) ? __ldg(ab + (size_t)m * HWa + q + i) : 0.f;
      }
      op_store8<TERMS>(a_hi, a_lo, r, k8, v);
    }
    // ---- B operand: rows = (cb, ky, kx) n0.., k = pixels
    for (int task = tid; task < BN * (KC / 8); task += PK_THREADS) {
      const int r = task >> 2, k8 = task & 3;
      const int n = n0 + r, q = q0 + k8 * 8;
      float v[8];
      if (n >= Ntot) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      } else if (!GENERAL) {
        // 1x1: n is the channel; contiguous pixels (HW % 8 == 0 guaranteed by the launcher)
        const float* sp = (n < p.CB1) ? p.b + (size_t)b * p.b_bs + (size_t)(g * p.CB1 + n) * HWb
                                      : p.b2 + (size_t)b * p.b2_bs + (size_t)(n - p.CB1) * HWb;
        if (q + 8 <= HWa) {
          const float4* src = reinterpret_cast<const float4*>(sp + q);
          const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
          v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
          v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
          if (LN) {
            const float4* stp = reinterpret_cast<const float4*>(p.ln_stats + ((size_t)b * HWb + q) * 2);
            const float ga = __ldg(p.ln_gamma + n), be = __ldg(p.ln_beta + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 s2 = __ldg(stp + j);  // (mean, rstd) of two pixels
              v[2 * j] = (v[2 * j] - s2.x) * s2.y * ga + be;
              v[2 * j + 1] = (v[2 * j + 1] - s2.z) * s2.w * ga + be;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = 0.f;
            if (q + i < HWa) {
              x = __ldg(sp + q + i);
              if (LN) {
                const float2 s2 = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)b * HWb + q + i);
                x = (x - s2.x) * s2.y * __ldg(p.ln_gamma + n) + __ldg(p.ln_beta + n);
              }
            }
            v[i] = x;
          }
        }
      } else {
        const int cb = n / KK, rr = n - cb * KK;
        const int ky = rr / p.ks, kx = rr - ky * p.ks;
        const float* sp = (cb < p.CB1) ? p.b + (size_t)b * p.b_bs + (size_t)(g * p.CB1 + cb) * HWb
                                       : p.b2 + (size_t)b * p.b2_bs + (size_t)(cb - p.CB1) * HWb;
        int qy = q / p.Wa, qx = q - qy * p.Wa;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = 0.f;
          if (q + i < HWa) {
            const int sy = qy * p.stride + ky - p.pad, sx = qx * p.stride + kx - p.pad;
            if ((unsigned)sy < (unsigned)p.Hb && (unsigned)sx < (unsigned)p.Wb) x = __ldg(sp + sy * p.Wb + sx);
          }
          v[i] = x;
          if (++qx == p.Wa) {
            qx = 0;
            ++qy;
          }
        }
      }
      op_store8<TERMS>(b_hi, b_lo, r, k8, v);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_stage<TERMS>(tmem, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), idesc, it == 0);
      tc_commit(&empty_bar[s]);
    }
  }
  if (c_end > c_begin) {
    if (tid == 0) tc_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    // ---- epilogue: thread = row m (TMEM lane), the two warp groups split the columns
    const uint32_t lane_base = tmem_lane_base(tmem);
    const int m = m0 + (warp & 3) * 32 + (tid & 31);
    const int half = warp >> 2;
    const int ncols8 = BN / 8;
    const int c8_begin = half ? ncols8 / 2 : 0, c8_end = half ? ncols8 : ncols8 / 2;
    float* ob = p.out + (size_t)g * p.out_gs + (p.per_image ? (size_t)bz * p.out_bs : 0);
    for (int c8 = c8_begin; c8 < c8_end; ++c8) {
      if (n0 + c8 * 8 >= Ntot) break;
      float v[8];
      tmem_ld8(lane_base + c8 * 8, v);
      if (m < p.CA) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = n0 + c8 * 8 + i;
          if (n < Ntot) atomicAdd(ob + (size_t)m * p.ldo + n, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <int TERMS, bool GENERAL, bool LN>
static int launch_pk(const rcot_pk_params& p, cudaStream_t stream) {
  const int Ntot = (p.CB1 + p.CB2) * p.ks * p.ks;
  const int HWa = p.Ha * p.Wa;
  int BN = round_up(Ntot, 16);
  if (BN > 256) BN = 256;
  const int nt = cdiv(Ntot, BN), mt = cdiv(p.CA, 128);
  const int cpi = cdiv(HWa, KC);
  const int total_chunks = p.per_image ? cpi : cpi * p.B;
  const int zdim = p.groups * (p.per_image ? p.B : 1);
  const long tiles = (long)mt * nt * zdim;
  int S = (int)((2 * 148 + tiles - 1) / tiles);
  int maxS = total_chunks / 4;
  if (maxS < 1) maxS = 1;
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  int per_cta = cdiv(total_chunks, S);
  S = cdiv(total_chunks, per_cta);
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  const size_t smem = (size_t)PK_STAGES * TA * (128 * KC * 2 + (size_t)BN * KC * 2);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pk_gemm_kernel<TERMS, GENERAL, LN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) {
      set_error("pk_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  RCOT_REQUIRE(zdim <= 65535 && S <= 65535, "pk_gemm: grid too large");
  dim3 grid(mt * nt, S, zdim);
  pk_gemm_kernel<TERMS, GENERAL, LN><<<grid, PK_THREADS, smem, stream>>>(p, BN, nt, cpi, per_cta, total_chunks,
                                                                        tmem_cols_pow2(BN));
  return check_launch("pk_gemm");
}

}  // namespace rcot

extern "C" int rcot_pk_gemm(const rcot_pk_params* pp, rcot_stream_t stream_) {
  using namespace rcot;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RCOT_REQUIRE(pp != nullptr, "pk_gemm: null params");
  rcot_pk_params p = *pp;
  RCOT_REQUIRE(p.a && p.b && p.out, "pk_gemm: null tensor pointer");
  RCOT_REQUIRE(p.B > 0 && p.CA > 0 && p.CB1 > 0 && p.CB2 >= 0, "pk_gemm: bad sizes");
  RCOT_REQUIRE((p.CB2 == 0) == (p.b2 == nullptr), "pk_gemm: b2/CB2 mismatch");
  RCOT_REQUIRE(p.Ha > 0 && p.Wa > 0 && p.Hb > 0 && p.Wb > 0 && p.ks >= 1 && p.stride >= 1, "pk_gemm: bad geometry");
  RCOT_REQUIRE(p.terms == 1 || p.terms == 3, "pk_gemm: terms must be 1 or 3");
  if (p.groups < 1) p.groups = 1;
  RCOT_REQUIRE(p.groups == 1 || p.CB2 == 0, "pk_gemm: groups and concat are exclusive");
  const bool ln = p.ln_stats != nullptr;
  const int HWa = p.Ha * p.Wa;
  bool general = !(p.ks == 1 && p.stride == 1 && p.pad == 0 && HWa % 8 == 0 && p.a_bs % 4 == 0 && p.b_bs % 4 == 0 &&
                   (p.b2 == nullptr || p.b2_bs % 4 == 0) && ((uintptr_t)p.a % 16 == 0) && ((uintptr_t)p.b % 16 == 0) &&
                   (p.b2 == nullptr || (uintptr_t)p.b2 % 16 == 0));
  if (p.ks == 1) RCOT_REQUIRE(p.Ha == p.Hb && p.Wa == p.Wb && p.stride == 1 && p.pad == 0, "pk_gemm: 1x1 geometry");
  if (ln) {
    RCOT_REQUIRE(p.ks == 1 && p.ln_gamma && p.ln_beta && p.groups == 1, "pk_gemm: LayerNorm needs ks==1");
    RCOT_REQUIRE(!general, "pk_gemm: LayerNorm path needs HW %% 8 == 0 and 16-byte aligned tensors");
    return p.terms == 3 ? launch_pk<3, false, true>(p, stream) : launch_pk<1, false, true>(p, stream);
  }
  if (general) return p.terms == 3 ? launch_pk<3, true, false>(p, stream) : launch_pk<1, true, false>(p, stream);
  return p.terms == 3 ? launch_pk<3, false, false>(p, stream) : launch_pk<1, false, false>(p, stream);
}
