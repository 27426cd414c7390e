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
#include "tmap.cuh"
#include <stdlib.h>

namespace rcot {

constexpr int PK_PROD_WARPS = 16;
constexpr int PK_PROD_THREADS = PK_PROD_WARPS * 32;
constexpr int PK_THREADS = PK_PROD_THREADS + 32;   // + one MMA-issuing warp
constexpr int PK_MAX_STAGES = 4;
// NBT (template): B-operand row tasks per producer thread, 1 for BN <= 128 and 2 for BN <= 256, so that narrow
// B operands (the common C = 48 / 96 cases) carry no dead second task through the conversion code.

// Producer thread t (16 warps) owns k8 = t & 3 (8 consecutive pixels of the 32-pixel chunk), A row t >> 2
// and B rows (t >> 2) + 128*j: four neighbouring lanes read 128 contiguous bytes of one channel row
// (the padded LBO of the operand layout keeps the matching shared-memory stores conflict free).
// The loads of chunk i+1 are issued into registers before chunk i is converted (bf16 hi/lo split)
// and stored, so global latency overlaps the conversion; one MMA warp issues tcgen05.mma and hands
// stages back through mbarriers; after the K loop the producer warps drain the TMEM accumulator
// with fp32 atomics (split-K).
template <int TERMS, bool GENERAL, bool LN, int NBT>
__global__ void __launch_bounds__(PK_THREADS, 1)
    pk_gemm_kernel(const rcot_pk_params p, const int BN, const int nt, const int cpi, const int per_cta,
                   const int total_chunks, const int stages, const uint32_t tmem_cols, const int tr) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[PK_MAX_STAGES], empty_bar[PK_MAX_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt_i = blockIdx.x / nt, nt_i = blockIdx.x - mt_i * nt;
  const int m0 = mt_i * 128, n0 = nt_i * BN;
  const int g = blockIdx.z % p.groups;
  const int bz = blockIdx.z / p.groups;  // image index when per_image, else 0
  const int HWa = p.Ha * p.Wa, HWb = p.Hb * p.Wb;
  const int KK = p.ks * p.ks;
  const int Ntot = (p.CB1 + p.CB2) * KK;

  const uint32_t a_tile = op_tile_bytes(128), b_tile = op_tile_bytes(BN);
  const uint32_t stage_bytes = TA * (a_tile + b_tile);

  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], PK_PROD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  const int c_begin = blockIdx.y * per_cta;
  int c_end = c_begin + per_cta;
  if (c_end > total_chunks) c_end = total_chunks;
  const int nchunks = c_end - c_begin;

  if (warp < PK_PROD_WARPS) {
    const int k8 = tid & 3, r0 = tid >> 2;   // 4 neighbouring lanes read 128 contiguous bytes of one row
    constexpr int nbt = NBT;                  // B row tasks (the launcher picks NBT = ceil(BN / 128))
    struct Regs {
      float a[8];
      float b[NBT][8];
      float2 st;                              // LayerNorm (mean, rstd) of pixel q0 + lane
    };
    // ---- everything that does not depend on the chunk is hoisted out of the K loop
    const bool a_ok = (m0 + r0) < p.CA;
    const float* a_row = p.a + (size_t)(g * p.CA + (a_ok ? m0 + r0 : 0)) * HWa + k8 * 8;
    bool b_ok[NBT];
    const float* b_row[NBT];
    int64_t b_bstride[NBT];
    float ga[NBT], be[NBT];
    int b_cb[NBT], b_ky[NBT], b_kx[NBT];
#pragma unroll
    for (int j = 0; j < NBT; ++j) {
      const int r = r0 + 128 * j, n = n0 + r;
      b_ok[j] = (j < nbt) && (r < BN) && (n < Ntot);
      const int cb = b_ok[j] ? n / KK : 0, rr = b_ok[j] ? n - cb * KK : 0;
      b_cb[j] = cb;
      b_ky[j] = rr / p.ks;
      b_kx[j] = rr - b_ky[j] * p.ks;
      const bool second = cb >= p.CB1;
      b_row[j] = second ? p.b2 + (size_t)(cb - p.CB1) * HWb : p.b + (size_t)(g * p.CB1 + cb) * HWb;
      b_bstride[j] = second ? p.b2_bs : p.b_bs;
      ga[j] = (LN && b_ok[j]) ? __ldg(p.ln_gamma + n) : 0.f;
      be[j] = (LN && b_ok[j]) ? __ldg(p.ln_beta + n) : 0.f;
    }
    const bool all_full = (HWa % KC) == 0;    // every chunk has 32 valid pixels
    auto ld8 = [&](float* v, const float* src) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(src)), x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
      v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    };
    auto load = [&](Regs& R, int b, int q0) {
      const int q = q0 + k8 * 8;
      if (!GENERAL && all_full) {
        // fast path: aligned 32-byte reads, no per-element predicates
        if (a_ok) ld8(R.a, a_row + (size_t)b * p.a_bs + q0);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) R.a[i] = 0.f;
        }
        if (LN) R.st = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)b * HWb + q0 + lane);
#pragma unroll
        for (int j = 0; j < NBT; ++j) {
          if (b_ok[j]) ld8(R.b[j], b_row[j] + (size_t)b * (LN ? p.b_bs : b_bstride[j]) + q);
          else {
#pragma unroll
            for (int i = 0; i < 8; ++i) R.b[j][i] = 0.f;
          }
        }
        return;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        R.a[i] = (a_ok && q + i < HWa) ? __ldg(a_row + (size_t)b * p.a_bs + q0 + i) : 0.f;
      if (LN) {
        const int ql = q0 + lane;
        R.st = ql < HWa ? __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)b * HWb + ql)
                        : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < NBT; ++j) {
        const float* sp = b_row[j] + (size_t)b * b_bstride[j];
        if (!b_ok[j]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) R.b[j][i] = 0.f;
        } else if (!GENERAL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) R.b[j][i] = (q + i < HWa) ? __ldg(sp + q + i) : 0.f;
        } else if (p.stride == 1 && (p.Wa & 7) == 0 && q + 8 <= HWa) {
          // stride-1 conv on a map whose width is a multiple of 8: this thread's 8 pixels share one image row, so
          // the row test is done once and only the two ends of the 8-float segment can fall outside
          const int qy = q / p.Wa, qx = q - qy * p.Wa;
          const int sy = qy + b_ky[j] - p.pad, sx0 = qx + b_kx[j] - p.pad;
          const bool row_in = (unsigned)sy < (unsigned)p.Hb;
          const float* r = sp + (row_in ? sy : 0) * p.Wb + sx0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            R.b[j][i] = (row_in && (unsigned)(sx0 + i) < (unsigned)p.Wb) ? __ldg(r + i) : 0.f;
        } else {
          int qy = q / p.Wa, qx = q - qy * p.Wa;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = 0.f;
            if (q + i < HWa) {
              const int sy = qy * p.stride + b_ky[j] - p.pad, sx = qx * p.stride + b_kx[j] - p.pad;
              if ((unsigned)sy < (unsigned)p.Hb && (unsigned)sx < (unsigned)p.Wb) x = __ldg(sp + sy * p.Wb + sx);
            }
            R.b[j][i] = x;
            if (++qx == p.Wa) {
              qx = 0;
              ++qy;
            }
          }
        }
      }
    };
    // (image, first pixel) of the chunk being prefetched, advanced without divisions
    int nb = p.per_image ? bz : c_begin / cpi;
    int nq0 = (p.per_image ? c_begin : c_begin - nb * cpi) * KC;
    int s = 0;
    uint32_t ph = 0;
    auto advance = [&]() {            // walker -> next chunk
      nq0 += KC;
      if (!p.per_image && nq0 >= cpi * KC) {
        nq0 = 0;
        ++nb;
      }
    };
    auto process = [&](Regs& cur, int cq0) {   // convert + store one prefetched chunk into stage s
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      uint8_t* a_hi = st;
      uint8_t* a_lo = st + a_tile;
      uint8_t* b_hi = st + TA * a_tile;
      uint8_t* b_lo = b_hi + b_tile;
      op_store8<TERMS>(a_hi, a_lo, r0, k8, cur.a);
      if (LN) {
        // statistics of this thread's 8 pixels are held by lanes k8*8+i of the warp (all lanes shuffle)
        const int q = cq0 + k8 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float mu = __shfl_sync(0xffffffffu, cur.st.x, k8 * 8 + i);
          const float rv = __shfl_sync(0xffffffffu, cur.st.y, k8 * 8 + i);
          const bool in = all_full || q + i < HWa;
#pragma unroll
          for (int j = 0; j < NBT; ++j)
            cur.b[j][i] = (in && b_ok[j]) ? (cur.b[j][i] - mu) * rv * ga[j] + be[j] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < NBT; ++j) {
        if (j >= nbt) break;
        const int r = r0 + 128 * j;
        if (r < BN) op_store8<TERMS>(b_hi, b_lo, r, k8, cur.b[j]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    };
    if (LN && NBT > 1) {
      // (the wide LayerNorm variant is register-bound at 544 threads: one prefetch set + a copy is the cheaper shape)
      Regs nxt;
      int qn = nq0;
      if (nchunks > 0) {
        load(nxt, nb, nq0);
        advance();
      }
      for (int it = 0; it < nchunks; ++it) {
        Regs cur = nxt;
        const int qc = qn;
        if (it + 1 < nchunks) {
          qn = nq0;
          load(nxt, nb, nq0);
          advance();
        }
        process(cur, qc);
      }
    } else {
      // two register sets in ping-pong: the loads of chunk i+1 fly while chunk i is converted, no copies
      Regs ra, rb;
      int qa = nq0, qb = 0;
      if (nchunks > 0) {
        load(ra, nb, nq0);
        advance();
      }
      for (int it = 0; it < nchunks; it += 2) {
        const bool has_b = it + 1 < nchunks;
        if (has_b) {
          qb = nq0;
          load(rb, nb, nq0);
          advance();
        }
        process(ra, qa);
        if (has_b) {
          if (it + 2 < nchunks) {
            qa = nq0;
            load(ra, nb, nq0);
            advance();
          }
          process(rb, qb);
        }
      }
    }
  } else {
    // ---- MMA issuer warp
    const uint32_t idesc = make_idesc_bf16(128, BN);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nchunks; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        issue_stage<TERMS>(tmem, st, st + a_tile, st + TA * a_tile, st + TA * a_tile + b_tile, idesc, it == 0);
        tc_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
    if (elect_one() && nchunks > 0) tc_commit(&done_bar);
  }
  if (warp < PK_PROD_WARPS && nchunks > 0) {
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    // ---- epilogue: thread = row m (TMEM lane), the four warp groups split the columns.
    // tr = 1: the launcher swapped the operands (see rcot_pk_gemm), so kernel row m is the caller's COLUMN: the
    // 32 lanes of a warp then add to consecutive addresses (coalesced reductions).  Otherwise 16-byte vector
    // reductions when the output rows are 16-byte aligned, scalar atomics as the last resort.
    const uint32_t lane_base = tmem_lane_base(tmem);
    const int m = m0 + (warp & 3) * 32 + lane;
    const int part = warp >> 2;               // 4 column parts
    const int ncols8 = BN / 8;
    const int c8_begin = (ncols8 * part) / 4, c8_end = (ncols8 * (part + 1)) / 4;
    float* ob = p.out + (size_t)g * p.out_gs + (p.per_image ? (size_t)bz * p.out_bs : 0);
    const bool vec = !tr && ((reinterpret_cast<uintptr_t>(ob) & 15) == 0) && (p.ldo % 4 == 0);
    for (int c8 = c8_begin; c8 < c8_end; ++c8) {
      if (n0 + c8 * 8 >= Ntot) break;
      float v[8];
      tmem_ld8(lane_base + c8 * 8, v);
      if (m < p.CA) {
        if (tr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = n0 + c8 * 8 + i;
            if (n < Ntot) atomicAdd(ob + (size_t)n * p.ldo + m, v[i]);
          }
        } else if (vec && n0 + c8 * 8 + 8 <= Ntot) {
          float* dst = ob + (size_t)m * p.ldo + n0 + c8 * 8;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                       "f"(v[3])
                       : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]),
                       "f"(v[7])
                       : "memory");
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = n0 + c8 * 8 + i;
            if (n < Ntot) atomicAdd(ob + (size_t)m * p.ldo + n, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-M variant for the blocks' 1x1 weight gradients with LayerNorm (dW_in = du z^T, dW_qkv = dpre z^T): CA is
// 2.6..5.3 x CB there, so a CTA that owns ONE 128-row tile of A re-reads and re-normalises the same B chunk for
// every tile.  Here a CTA owns MT (2 or 4) A tiles = MT accumulators in TMEM (MT*BN <= 512 columns): the B operand
// (x, normalised on the fly) is converted once per 32-pixel chunk for all of them.  One register set per thread:
// each operand row is stored and immediately refilled with the next chunk's loads, so global latency hides behind
// the rest of the chunk.  Rows beyond CA / columns beyond N are simply never written or read back (an accumulator
// element depends on one A row and one B row only).
template <int TERMS, int MT, int NBT, int ABF = 0>
__global__ void __launch_bounds__(PK_THREADS, 1)
    pk_mm_kernel(const rcot_pk_params p, const int BN, const int nt, const int cpi, const int per_cta,
                 const int total_chunks, const int stages, const uint32_t tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[PK_MAX_STAGES], empty_bar[PK_MAX_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mg = blockIdx.x / nt, nt_i = blockIdx.x - mg * nt;
  const int m0 = mg * (MT * 128), n0 = nt_i * BN;
  const int HW = p.Ha * p.Wa;
  const int Ntot = p.CB1;
  const uint32_t a_tile = op_tile_bytes(128), b_tile = op_tile_bytes(BN);
  constexpr int TAA = ABF ? 1 : TA;           // a bf16-stored A operand is exact: no lo image
  const uint32_t stage_bytes = TAA * MT * a_tile + TA * b_tile;
  int mt_valid = (p.CA - m0 + 127) >> 7;      // A tiles of this CTA that hold at least one row
  if (mt_valid > MT) mt_valid = MT;

  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], PK_PROD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  const int c_begin = blockIdx.y * per_cta;
  int c_end = c_begin + per_cta;
  if (c_end > total_chunks) c_end = total_chunks;
  const int nchunks = c_end - c_begin;

  if (warp < PK_PROD_WARPS) {
    const int k8 = tid & 3, r0 = tid >> 2;   // 4 neighbouring lanes read 128 contiguous bytes of one row
    // Loads are unconditional (rows beyond CA / N are clamped to the last valid row and the walker stops at the
    // CTA's last chunk): a load whose result is merged under a predicate makes the compiler wait for it right
    // away, which exposes the full global latency every chunk (measured: 80 % long-scoreboard stalls).
    const float* a_ptr[MT];                  // element offsets are the same for a bf16 A: indexed as 2-byte below
    bool a_ok[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int row = m0 + m * 128 + r0;
      a_ok[m] = row < p.CA;
      a_ptr[m] = p.a + (size_t)(a_ok[m] ? row : p.CA - 1) * HW + k8 * 8;
    }
    const __nv_bfloat16* a16 = reinterpret_cast<const __nv_bfloat16*>(p.a);
    auto ld8a = [&](uint4& v16, float* v, const float* src) {
      if (ABF) v16 = __ldg(reinterpret_cast<const uint4*>(a16 + (src - p.a)));   // 8 pixels = 16 bytes, already the operand
      else {
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(src)), x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
        v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
      }
    };
    uint4 ra16[MT];
    bool b_ok[NBT];
    const float* b_ptr[NBT];
    float ga[NBT], be[NBT];
#pragma unroll
    for (int j = 0; j < NBT; ++j) {
      const int r = r0 + 128 * j, n = n0 + r;
      b_ok[j] = (r < BN) && (n < Ntot);
      const int nc = b_ok[j] ? n : Ntot - 1;
      b_ptr[j] = p.b + (size_t)nc * HW + k8 * 8;
      ga[j] = __ldg(p.ln_gamma + nc);
      be[j] = __ldg(p.ln_beta + nc);
    }
    const float2* stats = reinterpret_cast<const float2*>(p.ln_stats);
    auto ld8 = [&](float* v, const float* src) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(src)), x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
      v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    };
    float ra[MT][8], rb[NBT][8];
    // (image, first pixel) of the chunk whose loads are issued next
    int nb = c_begin / cpi;
    int nq0 = (c_begin - nb * cpi) * KC;
    {
      const size_t ao = (size_t)nb * p.a_bs + nq0, bo = (size_t)nb * p.b_bs + nq0;
#pragma unroll
      for (int m = 0; m < MT; ++m) ld8a(ra16[m], ra[m], a_ptr[m] + ao);
#pragma unroll
      for (int j = 0; j < NBT; ++j) ld8(rb[j], b_ptr[j] + bo);
    }
    float2 st = __ldg(stats + (size_t)nb * HW + nq0 + lane);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nchunks; ++it) {
      if (it + 1 < nchunks) {                     // walker -> the chunk after this one (stays on the last chunk)
        nq0 += KC;
        if (nq0 >= HW) {
          nq0 = 0;
          ++nb;
        }
      }
      const size_t ao = (size_t)nb * p.a_bs + nq0, bo = (size_t)nb * p.b_bs + nq0;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* stg = smem + (size_t)s * stage_bytes;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (a_ok[m]) {
          if (ABF) {
            const uint32_t pk4[4] = {ra16[m].x, ra16[m].y, ra16[m].z, ra16[m].w};
            op_store8_bf16(stg + m * a_tile, r0, k8, pk4);
          } else {
            op_store8<TERMS>(stg + m * a_tile, stg + (MT + m) * a_tile, r0, k8, ra[m]);
          }
        }
        ld8a(ra16[m], ra[m], a_ptr[m] + ao);
      }
      const float2 stc = st;                      // statistics of pixel q0 + lane of THIS chunk
      st = __ldg(stats + (size_t)nb * HW + nq0 + lane);   // (requesting these earlier in the iteration measured slower)
#pragma unroll
      for (int i = 0; i < 8; ++i) {               // this thread's 8 pixels are held by lanes k8*8+i (all lanes shuffle)
        const float mu = __shfl_sync(0xffffffffu, stc.x, k8 * 8 + i);
        const float rv = __shfl_sync(0xffffffffu, stc.y, k8 * 8 + i);
#pragma unroll
        for (int j = 0; j < NBT; ++j) rb[j][i] = fmaf((rb[j][i] - mu) * rv, ga[j], be[j]);
      }
      uint8_t* b_hi = stg + TAA * MT * a_tile;
#pragma unroll
      for (int j = 0; j < NBT; ++j) {
        if (b_ok[j]) op_store8<TERMS>(b_hi, b_hi + b_tile, r0 + 128 * j, k8, rb[j]);
        ld8(rb[j], b_ptr[j] + bo);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    // ---- MMA issuer warp: MT accumulators share the B stage
    const uint32_t idesc = make_idesc_bf16(128, BN);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nchunks; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t stg = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_hi = stg + TAA * MT * a_tile;
#pragma unroll
        for (int m = 0; m < MT; ++m)
          if (m < mt_valid)
            issue_stage<TERMS, !ABF>(tmem + m * BN, stg + m * a_tile, stg + (MT + m) * a_tile, b_hi, b_hi + b_tile, idesc,
                                     it == 0);
        tc_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
    if (elect_one() && nchunks > 0) tc_commit(&done_bar);
  }
  if (warp < PK_PROD_WARPS && nchunks > 0) {
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    // ---- epilogue: thread = row (TMEM lane), the four warp groups split the columns; 16-byte vector reductions
    // (red.global.add.v4.f32) when the output rows are 16-byte aligned, scalar atomics otherwise
    const uint32_t lane_base = tmem_lane_base(tmem);
    const int part = warp >> 2;
    const int ncols8 = BN / 8;
    const int c8_begin = (ncols8 * part) / 4, c8_end = (ncols8 * (part + 1)) / 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (p.ldo % 4 == 0) && (n0 % 4 == 0);
    for (int m = 0; m < mt_valid; ++m) {
      const int row = m0 + m * 128 + (warp & 3) * 32 + lane;
      for (int c8 = c8_begin; c8 < c8_end; ++c8) {
        if (n0 + c8 * 8 >= Ntot) break;
        float v[8];
        tmem_ld8(lane_base + m * BN + c8 * 8, v);
        if (row < p.CA) {
          float* dst = p.out + (size_t)row * p.ldo + n0 + c8 * 8;
          if (vec && n0 + c8 * 8 + 8 <= Ntot) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                         "f"(v[3])
                         : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]),
                         "f"(v[6]), "f"(v[7])
                         : "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (n0 + c8 * 8 + i < Ntot) atomicAdd(dst + i, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}


// Split-K factor.  These kernels fill an SM's shared memory, so their CTAs are resident ONE per SM and a grid runs in
// rounds of 148: 297 CTAs (what "two per SM, rounded up" gave for the 1x1 weight gradients with three tiles) take three
// rounds, the last one for a single CTA -- ncu showed such launches at 2.7 TB/s.  Cost model: rounds x (chunks per CTA +
// a fixed per-CTA cost of ~8 chunk times for TMEM allocation, pipeline fill and the atomic epilogue); the smallest
// cost wins.  RCOT_PK_SPLIT=0 restores the round-1 rule (A/B switch).
static int pick_split(long tiles, int total_chunks, int maxS, int old_target) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("RCOT_PK_SPLIT");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (maxS < 1) maxS = 1;
  if (!enabled) {
    int S = (int)((old_target + tiles - 1) / tiles);
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
    return cdiv(total_chunks, cdiv(total_chunks, S));
  }
  const int sms = 148, fixed = 8;
  long best_cost = -1;
  int best = 1;
  for (int S = 1; S <= maxS && (S == 1 || tiles * S <= 4L * sms); ++S) {
    const int per = cdiv(total_chunks, S);
    const int Se = cdiv(total_chunks, per);
    const long rounds = (tiles * Se + sms - 1) / sms;
    const long cost = rounds * (per + fixed);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = Se;
    }
  }
  return best;
}

template <int TERMS, int MT, int NBT, int ABF = 0>
static int launch_pk_mm_t(const rcot_pk_params& p, int BN, cudaStream_t stream) {
  const int Ntot = p.CB1;
  const int HW = p.Ha * p.Wa;
  const int nt = cdiv(Ntot, BN), mgroups = cdiv(p.CA, 128 * MT);
  const int cpi = HW / KC;
  const int total_chunks = cpi * p.B;
  const long tiles = (long)mgroups * nt;
  int S = pick_split(tiles, total_chunks, total_chunks / 4, 148);   // every split adds CA x N atomics to the epilogue
  int per_cta = cdiv(total_chunks, S);
  S = cdiv(total_chunks, per_cta);
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  constexpr int TAA = ABF ? 1 : TA;
  const size_t stage_bytes = (size_t)TAA * MT * (size_t)op_tile_bytes(128) + (size_t)TA * (size_t)op_tile_bytes(BN);
  int stages = (int)((196 * 1024) / stage_bytes);
  if (stages > PK_MAX_STAGES) stages = PK_MAX_STAGES;
  RCOT_REQUIRE(stages >= 2, "pk_gemm(mm): stage of %zu bytes does not fit twice", stage_bytes);
  const size_t smem = stages * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pk_mm_kernel<TERMS, MT, NBT, ABF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         196 * 1024);
    if (e != cudaSuccess) {
      set_error("pk_gemm(mm): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  RCOT_REQUIRE(S <= 65535, "pk_gemm(mm): grid too large");
  dim3 grid(mgroups * nt, S, 1);
  pk_mm_kernel<TERMS, MT, NBT, ABF><<<grid, PK_THREADS, smem, stream>>>(p, BN, nt, cpi, per_cta, total_chunks, stages,
                                                                       tmem_cols_pow2(MT * BN));
  return check_launch("pk_gemm(mm)");
}
template <int TERMS, int MT, int NBT>
static int launch_pk_mm(const rcot_pk_params& p, int BN, cudaStream_t stream) {
  return p.a_bf16 ? launch_pk_mm_t<TERMS, MT, NBT, 1>(p, BN, stream) : launch_pk_mm_t<TERMS, MT, NBT, 0>(p, BN, stream);
}

// Picks (MT, NBT) for the multi-M kernel; returns -100 when the call does not qualify.
template <int TERMS>
static int try_pk_mm(const rcot_pk_params& p, cudaStream_t stream) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("RCOT_PK_MM");       // RCOT_PK_MM=0: A/B switch back to the one-tile-per-CTA kernel
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const int HW = p.Ha * p.Wa;
  if (!enabled || p.CA <= 128 || HW % KC != 0 || p.per_image || p.groups > 1 || p.b2 != nullptr) return -100;
  int BN = round_up(p.CB1, 16);
  if (BN > 256) BN = 256;
  if (BN <= 128) {
    if (p.CA > 256) return launch_pk_mm<TERMS, 4, 1>(p, BN, stream);
    return launch_pk_mm<TERMS, 2, 1>(p, BN, stream);
  }
  return launch_pk_mm<TERMS, 2, 2>(p, BN, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-staged variant of the plain 1x1 kernel (per-image Grams q k^T / dy v^T, dW of GDFN's project_out): both
// operands are K-major in NCHW as they lie, so a 32-pixel chunk of the A tile (128 channels) and of the B tile
// (BN <= 128 channels) is ONE 3-D tensor copy each (box 32 px x rows x 1 image) into a ring of raw fp32 slots issued
// by a loader warp; the 16 producer warps convert from shared memory.  With register prefetch one chunk ahead the
// producers keep ~40 KB in flight per SM (3.1-3.4 TB/s measured); the ring holds PK_RAW chunks.
constexpr int PK_RAW_MAX = 4;

template <int TERMS, int ABF = 0, int BBF = 0>
__global__ void __launch_bounds__(PK_THREADS + 32, 1)
    pk_tma_kernel(const rcot_pk_params p, const int BN, const int nt, const int cpi, const int per_cta,
                  const int total_chunks, const int stages, const int nraw, const uint32_t tmem_cols, const int tr,
                  const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[PK_MAX_STAGES], empty_bar[PK_MAX_STAGES], done_bar;
  __shared__ uint64_t raw_full[PK_RAW_MAX], raw_empty[PK_RAW_MAX];
  __shared__ uint32_t tmem_base_s;
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt_i = blockIdx.x / nt, nt_i = blockIdx.x - mt_i * nt;
  const int m0 = mt_i * 128, n0 = nt_i * BN;
  const int g = blockIdx.z % p.groups;
  const int bz = blockIdx.z / p.groups;  // image index when per_image, else 0
  const int HW = p.Ha * p.Wa;
  const int Ntot = p.CB1;
  const uint32_t a_tile = op_tile_bytes(128), b_tile = op_tile_bytes(BN);
  constexpr int TAA = ABF ? 1 : TA, TBB = BBF ? 1 : TA;           // bf16-stored operands are exact: no lo image
  const uint32_t stage_bytes = TAA * a_tile + TBB * b_tile;
  const uint32_t rawA_bytes = 128u * KC * sizeof(float);           // 16 KB: 128 channels x 32 pixels (slot stride)
  const uint32_t raw_bytes = rawA_bytes + (uint32_t)BN * KC * sizeof(float);
  const uint32_t tx_bytes = 128u * KC * (ABF ? 2 : 4) + (uint32_t)BN * KC * (BBF ? 2 : 4);
  uint8_t* raw_ring = smem + (size_t)stages * stage_bytes;

  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], PK_PROD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < nraw; ++i) {
      mbar_init(&raw_full[i], 1);               // the loader's expect_tx arrival (+ the two copies' bytes)
      mbar_init(&raw_empty[i], PK_PROD_WARPS);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  const int c_begin = blockIdx.y * per_cta;
  int c_end = c_begin + per_cta;
  if (c_end > total_chunks) c_end = total_chunks;
  const int nchunks = c_end - c_begin;

  if (warp == PK_PROD_WARPS + 1) {
    // ---- loader: one tensor copy per operand and chunk
    int nb = p.per_image ? bz : c_begin / cpi;
    int nq0 = (p.per_image ? c_begin : c_begin - nb * cpi) * KC;
    int rs = 0;
    uint32_t rph = 0;
    for (int it = 0; it < nchunks; ++it) {
      mbar_wait(&raw_empty[rs], rph ^ 1);
      if (lane == 0) {
        uint8_t* slot = raw_ring + (size_t)rs * raw_bytes;
        mbar_arrive_expect_tx(&raw_full[rs], tx_bytes);        // rows outside the tensors are zero-filled and counted
        tensor_g2s_3d(slot, &tmA, nq0, g * p.CA + m0, nb, &raw_full[rs]);
        tensor_g2s_3d(slot + rawA_bytes, &tmB, nq0, g * p.CB1 + n0, nb, &raw_full[rs]);
      }
      __syncwarp();
      if (++rs == nraw) {
        rs = 0;
        rph ^= 1;
      }
      nq0 += KC;
      if (!p.per_image && nq0 >= HW) {
        nq0 = 0;
        ++nb;
      }
    }
  } else if (warp < PK_PROD_WARPS) {
    // ---- producers: raw slot -> bf16 hi/lo operand stage.  Thread = (row r0, 8-pixel group k8).
    const int k8 = tid & 3, r0 = tid >> 2;
    const bool a_ok = (m0 + r0) < p.CA;
    const bool b_ok = (r0 < BN) && (n0 + r0) < Ntot;
    int rs = 0, s = 0;
    uint32_t rph = 0, ph = 0;
    for (int it = 0; it < nchunks; ++it) {
      mbar_wait(&raw_full[rs], rph);
      const uint8_t* slot = raw_ring + (size_t)rs * raw_bytes;
      float va[8], vb[8];
      uint4 va16 = make_uint4(0, 0, 0, 0), vb16 = make_uint4(0, 0, 0, 0);
      if (ABF) {                        // rows of 32 bf16 pixels (64 bytes): this thread's 8 pixels are one 16-byte word
        va16 = *reinterpret_cast<const uint4*>(slot + (size_t)r0 * (KC * 2) + k8 * 16);
      } else {
        const float4* ra = reinterpret_cast<const float4*>(slot + (size_t)r0 * (KC * 4) + k8 * 32);
        const float4 x0 = ra[0], x1 = ra[1];
        va[0] = x0.x; va[1] = x0.y; va[2] = x0.z; va[3] = x0.w;
        va[4] = x1.x; va[5] = x1.y; va[6] = x1.z; va[7] = x1.w;
      }
      if (BBF) {
        vb16 = *reinterpret_cast<const uint4*>(slot + rawA_bytes + (size_t)(b_ok ? r0 : 0) * (KC * 2) + k8 * 16);
      } else {
        const float4* rb = reinterpret_cast<const float4*>(slot + rawA_bytes + (size_t)(b_ok ? r0 : 0) * (KC * 4) + k8 * 32);
        const float4 y0 = rb[0], y1 = rb[1];
        vb[0] = y0.x; vb[1] = y0.y; vb[2] = y0.z; vb[3] = y0.w;
        vb[4] = y1.x; vb[5] = y1.y; vb[6] = y1.z; vb[7] = y1.w;
      }
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      if (a_ok) {
        if (ABF) {
          const uint32_t pk4[4] = {va16.x, va16.y, va16.z, va16.w};
          op_store8_bf16(st, r0, k8, pk4);
        } else {
          op_store8<TERMS>(st, st + a_tile, r0, k8, va);
        }
      }
      if (b_ok) {
        if (BBF) {
          const uint32_t pk4[4] = {vb16.x, vb16.y, vb16.z, vb16.w};
          op_store8_bf16(st + TAA * a_tile, r0, k8, pk4);
        } else {
          op_store8<TERMS>(st + TAA * a_tile, st + TAA * a_tile + b_tile, r0, k8, vb);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full_bar[s]);
        mbar_arrive(&raw_empty[rs]);     // every lane's raw values have been consumed (stored) by now
      }
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
      if (++rs == nraw) {
        rs = 0;
        rph ^= 1;
      }
    }
  } else if (warp == PK_PROD_WARPS) {
    // ---- MMA issuer warp
    const uint32_t idesc = make_idesc_bf16(128, BN);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nchunks; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        issue_stage<TERMS, !ABF, !BBF>(tmem, st, st + a_tile, st + TAA * a_tile, st + TAA * a_tile + b_tile, idesc, it == 0);
        tc_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        ph ^= 1;
      }
    }
    if (elect_one() && nchunks > 0) tc_commit(&done_bar);
  }
  if (warp < PK_PROD_WARPS && nchunks > 0) {
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    // ---- epilogue (same as pk_gemm_kernel): thread = row m, four warp groups split the columns
    const uint32_t lane_base = tmem_lane_base(tmem);
    const int m = m0 + (warp & 3) * 32 + lane;
    const int part = warp >> 2;
    const int ncols8 = BN / 8;
    const int c8_begin = (ncols8 * part) / 4, c8_end = (ncols8 * (part + 1)) / 4;
    float* ob = p.out + (size_t)g * p.out_gs + (p.per_image ? (size_t)bz * p.out_bs : 0);
    const bool vec = !tr && ((reinterpret_cast<uintptr_t>(ob) & 15) == 0) && (p.ldo % 4 == 0);
    for (int c8 = c8_begin; c8 < c8_end; ++c8) {
      if (n0 + c8 * 8 >= Ntot) break;
      float v[8];
      tmem_ld8(lane_base + c8 * 8, v);
      if (m < p.CA) {
        if (tr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = n0 + c8 * 8 + i;
            if (n < Ntot) atomicAdd(ob + (size_t)n * p.ldo + m, v[i]);
          }
        } else if (vec && n0 + c8 * 8 + 8 <= Ntot) {
          float* dst = ob + (size_t)m * p.ldo + n0 + c8 * 8;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                       "f"(v[3])
                       : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]),
                       "f"(v[7])
                       : "memory");
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = n0 + c8 * 8 + i;
            if (n < Ntot) atomicAdd(ob + (size_t)m * p.ldo + n, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// Returns -100 when the call does not qualify (the caller then launches the register-staged kernel).
template <int TERMS, int ABF = 0, int BBF = 0>
static int try_pk_tma_t(const rcot_pk_params& p, cudaStream_t stream, int tr) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("RCOT_PK_TMA");      // RCOT_PK_TMA=0: A/B switch back to the register-staged kernel
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const int HW = p.Ha * p.Wa;
  const int Ntot = p.CB1;
  int BN = round_up(Ntot, 16);
  if (BN > 256) BN = 256;
  if (!enabled || tensor_map_encoder() == nullptr || HW % KC != 0 || BN > 128 || p.b2 != nullptr) return -100;
  const int nt = cdiv(Ntot, BN), mt = cdiv(p.CA, 128);
  const int cpi = HW / KC;
  const int total_chunks = p.per_image ? cpi : cpi * p.B;
  const int zdim = p.groups * (p.per_image ? p.B : 1);
  const long tiles = (long)mt * nt * zdim;
  int S = pick_split(tiles, total_chunks, total_chunks / 4, 2 * 148);
  int per_cta = cdiv(total_chunks, S);
  S = cdiv(total_chunks, per_cta);
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  const size_t stage_bytes = (size_t)(ABF ? 1 : TA) * op_tile_bytes(128) + (size_t)(BBF ? 1 : TA) * (size_t)op_tile_bytes(BN);
  const size_t raw_bytes = (size_t)(128 + BN) * KC * sizeof(float);
  int nraw = 3, stages = (int)((196 * 1024 - nraw * raw_bytes) / stage_bytes);
  if (stages > PK_MAX_STAGES) stages = PK_MAX_STAGES;
  if (stages > 3) {                              // room left: a fourth raw slot instead of a fourth operand stage
    nraw = 4;
    stages = (int)((196 * 1024 - nraw * raw_bytes) / stage_bytes);
    if (stages > 3) stages = 3;
  }
  if (stages < 2) return -100;
  const size_t smem = stages * stage_bytes + nraw * raw_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pk_tma_kernel<TERMS, ABF, BBF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196 * 1024);
    if (e != cudaSuccess) {
      set_error("pk_gemm(tma): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  RCOT_REQUIRE(zdim <= 65535 && S <= 65535, "pk_gemm(tma): grid too large");
  CUtensorMap tmA, tmB;
  int rc = make_act_map(&tmA, p.a, p.a_bs, p.groups * p.CA, HW, p.B, KC, 128, "pk_gemm", ABF);
  if (rc == RCOT_OK) rc = make_act_map(&tmB, p.b, p.b_bs, p.groups * p.CB1, HW, p.B, KC, BN, "pk_gemm", BBF);
  if (rc != RCOT_OK) return rc;
  dim3 grid(mt * nt, S, zdim);
  pk_tma_kernel<TERMS, ABF, BBF><<<grid, PK_THREADS + 32, smem, stream>>>(p, BN, nt, cpi, per_cta, total_chunks, stages,
                                                                        nraw, tmem_cols_pow2(BN), tr, tmA, tmB);
  return check_launch("pk_gemm(tma)");
}
template <int TERMS>
static int try_pk_tma(const rcot_pk_params& p, cudaStream_t stream, int tr) {
  if (p.a_bf16 && p.b_bf16) return try_pk_tma_t<TERMS, 1, 1>(p, stream, tr);
  if (p.a_bf16) return try_pk_tma_t<TERMS, 1, 0>(p, stream, tr);
  if (p.b_bf16) return try_pk_tma_t<TERMS, 0, 1>(p, stream, tr);
  return try_pk_tma_t<TERMS, 0, 0>(p, stream, tr);
}

template <int TERMS, bool GENERAL, bool LN, int NBT>
static int launch_pk_n(const rcot_pk_params& p, cudaStream_t stream, int tr) {
  const int Ntot = (p.CB1 + p.CB2) * p.ks * p.ks;
  const int HWa = p.Ha * p.Wa;
  int BN = round_up(Ntot, 16);
  if (BN > 256) BN = 256;
  const int nt = cdiv(Ntot, BN), mt = cdiv(p.CA, 128);
  const int cpi = cdiv(HWa, KC);
  const int total_chunks = p.per_image ? cpi : cpi * p.B;
  const int zdim = p.groups * (p.per_image ? p.B : 1);
  const long tiles = (long)mt * nt * zdim;
  int S = pick_split(tiles, total_chunks, total_chunks / 4, 2 * 148);
  int per_cta = cdiv(total_chunks, S);
  S = cdiv(total_chunks, per_cta);
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  const size_t stage_bytes = (size_t)TA * (op_tile_bytes(128) + (size_t)op_tile_bytes(BN));
  int stages = (int)((190 * 1024) / stage_bytes);
  if (stages > PK_MAX_STAGES) stages = PK_MAX_STAGES;
  const size_t smem = stages * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pk_gemm_kernel<TERMS, GENERAL, LN, NBT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 196 * 1024);
    if (e != cudaSuccess) {
      set_error("pk_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  RCOT_REQUIRE(zdim <= 65535 && S <= 65535, "pk_gemm: grid too large");
  dim3 grid(mt * nt, S, zdim);
  pk_gemm_kernel<TERMS, GENERAL, LN, NBT><<<grid, PK_THREADS, smem, stream>>>(p, BN, nt, cpi, per_cta, total_chunks,
                                                                             stages, tmem_cols_pow2(BN), tr);
  return check_launch("pk_gemm");
}

template <int TERMS, bool GENERAL, bool LN>
static int launch_pk(const rcot_pk_params& p, cudaStream_t stream, int tr = 0) {
  const int Ntot = (p.CB1 + p.CB2) * p.ks * p.ks;
  return Ntot <= 128 ? launch_pk_n<TERMS, GENERAL, LN, 1>(p, stream, tr)
                     : launch_pk_n<TERMS, GENERAL, LN, 2>(p, stream, tr);
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
  const bool any_bf16 = p.a_bf16 || p.b_bf16;
  if (any_bf16)
    RCOT_REQUIRE(p.ks == 1 && HWa % KC == 0 && p.b2 == nullptr && !(ln && p.b_bf16) && (p.a_bs % 8 == 0) && (p.b_bs % 8 == 0),
                 "pk_gemm: bf16 operands need a 1x1 product on H*W %% 32 == 0 without concat (LayerNorm input stays fp32)");
  bool general = !(p.ks == 1 && p.stride == 1 && p.pad == 0 && HWa % 8 == 0 && p.a_bs % 4 == 0 && p.b_bs % 4 == 0 &&
                   (p.b2 == nullptr || p.b2_bs % 4 == 0) && ((uintptr_t)p.a % 16 == 0) && ((uintptr_t)p.b % 16 == 0) &&
                   (p.b2 == nullptr || (uintptr_t)p.b2 % 16 == 0));
  if (p.ks == 1) RCOT_REQUIRE(p.Ha == p.Hb && p.Wa == p.Wb && p.stride == 1 && p.pad == 0, "pk_gemm: 1x1 geometry");
  if (ln) {
    RCOT_REQUIRE(p.ks == 1 && p.ln_gamma && p.ln_beta && p.groups == 1, "pk_gemm: LayerNorm needs ks==1");
    RCOT_REQUIRE(!general, "pk_gemm: LayerNorm path needs HW %% 8 == 0 and 16-byte aligned tensors");
    const int rc = p.terms == 3 ? try_pk_mm<3>(p, stream) : try_pk_mm<1>(p, stream);
    if (rc != -100) return rc;
    RCOT_REQUIRE(!any_bf16, "pk_gemm: a bf16 operand with LayerNorm needs the multi-M kernel (CA > 128)");
    return p.terms == 3 ? launch_pk<3, false, true>(p, stream) : launch_pk<1, false, true>(p, stream);
  }
  if (general) return p.terms == 3 ? launch_pk<3, true, false>(p, stream) : launch_pk<1, true, false>(p, stream);
  int tr = 0;
  if (p.b2 == nullptr && (p.ldo % 4 != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0)) {
    // The output rows are not 16-byte aligned (e.g. dW of GDFN's project_out: ldo = hid = 255): the two operands
    // of a 1x1 product are interchangeable, so swap them and let the epilogue write the transpose -- TMEM lanes
    // (consecutive threads) then map to consecutive output addresses and the reductions coalesce.
    rcot_pk_params q = p;
    q.a = p.b;
    q.a_bs = p.b_bs;
    q.CA = p.CB1;
    q.b = p.a;
    q.b_bs = p.a_bs;
    q.CB1 = p.CA;
    q.a_bf16 = p.b_bf16;
    q.b_bf16 = p.a_bf16;
    p = q;
    tr = 1;
  }
  {
    const int rc = p.terms == 3 ? try_pk_tma<3>(p, stream, tr) : try_pk_tma<1>(p, stream, tr);
    if (rc != -100) return rc;
  }
  RCOT_REQUIRE(!any_bf16, "pk_gemm: bf16 operands need the TMA-staged kernel (N <= 128 after the operand swap)");
  return p.terms == 3 ? launch_pk<3, false, false>(p, stream, tr) : launch_pk<1, false, false>(p, stream, tr);
}
