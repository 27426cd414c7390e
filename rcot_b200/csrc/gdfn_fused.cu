// gdfn_fused.cu -- GDFN forward as ONE kernel (Net_Restormer.py:80-85 with the block's LayerNorm and residual):
//     y = x + W_out . ( gelu(dw3x3(W_in . LN(x))[:hid]) * dw3x3(W_in . LN(x))[hid:] )
// The 5.3 C-wide hidden tensor u never reaches HBM: a CTA owns an 8 x 16-pixel tile of one image (+1-pixel halo = 180
// pixels = two 128-row MMA tiles) and walks the hidden dimension in SLICES of 16 (a, b) channel pairs:
//   Z   : LN(x) of the halo tile as the bf16 hi/lo A operand [256 x C], built once per tile and kept in TENSOR memory
//         (halo pixels outside the image become zero rows, so their u is the zero padding the depthwise conv expects);
//   per slice n:  GEMM-1  U_n[256 x 32] = Z . W_in[slice]^T          tcgen05 (A from TMEM) -> TMEM, 2 buffers
//                 drain   TMEM -> shared memory as [channel][10][18] fp32, 3 buffers        (4 DRAIN warps)
//                 stencil a = dw(U_n[:16]), b = dw(U_n[16:]), g = gelu(a) * b on the 128 core pixels (CUDA cores)
//                         -> bf16 hi/lo A operand [128 x 16], 2 slots                        (16 STENCIL warps)
//                 GEMM-2  Y[128 x C] += g . W_out[:, slice]^T         tcgen05, accumulates over the slices in TMEM
//   epilogue: Y (+ x) -> HBM, plus the per-pixel LayerNorm statistics the next block's LN1 needs.
// Version 5 pipeline (round 2): the phases of a slice are owned by DIFFERENT warps and linked by mbarriers only, so
// the TMEM drain, the shared-memory stores, the stencil loads, the FP32 pipe and the tensor pipe overlap instead of
// taking turns (v4 ran all 16 workers through drain -> barrier -> store -> barrier -> stencil in lockstep):
//   * one ISSUER warp (TMA weight rings + every tcgen05.mma) keeps GEMM-1 TWO slices ahead of GEMM-2;
//   * four DRAIN warps (one per TMEM lane quarter) move U_n to shared memory as soon as GEMM-1(n) retires;
//   * the sixteen STENCIL warps form two groups of eight that work on alternate slices (group = n & 1), each thread
//     a 2 x 4-pixel patch of one (a, b) pair: 4 halo rows are loaded once for two output rows and the 18 taps of
//     the pair once for 8 pixels -- 21 instead of 34 16-byte shared-memory loads per 8 outputs.
// Weight slices (pre-packed by rcot_gdfn_pack: operand images + the 32 x 9 depthwise taps) stream through the TMA
// engine (cp.async.bulk) into two 4-slot rings.  bf16x3 split products (hi*hi + lo*hi + hi*lo): fp32-class.
// Algorithmic HBM bytes: (1 + 180/128 halo re-read, mostly L2 hits) C + C per pixel instead of ~19 C for the three
// unfused launches.  Optional outputs u / g keep the existing (unfused) backward fed when it wants them saved.
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "gelu.cuh"
#include "fused_tile.cuh"
#include "tc.cuh"

namespace rcot {

constexpr int GF_S_WARPS = 16;                       // stencil / Z / epilogue warps: two groups of 8
constexpr int GF_D_WARPS = 4;                        // drain warps (TMEM lane quarter = warp & 3)
constexpr int GF_ISSUER = GF_S_WARPS + GF_D_WARPS;   // warp index of the GEMM-1 issuer; GF_ISSUER + 1 issues GEMM-2
constexpr int GF_WARPS = GF_ISSUER + 2;
constexpr int GF_THREADS = GF_WARPS * 32;            // 704
constexpr int GF_U_BUFS = 3;                         // shared-memory U slices in flight
constexpr int GF_W_SLOTS = 4;                        // W_in ring and W_out / tap ring
constexpr uint32_t GF_U_BYTES = 32 * GF_CS * 4;
constexpr uint32_t GF_G_SBO = 272;                   // g operand: 8-row group stride (256 + 16 B pad against bank conflicts)
constexpr uint32_t GF_G_TILE = 16 * GF_G_SBO;        // one term of the [128 x 16] g operand
constexpr uint32_t GF_DW_BYTES = 16 * 20 * sizeof(float);   // per pair: 9 a-taps, 9 b-taps, 2 pad (16-byte loads)

template <int C>
struct GfLayout {
  static constexpr uint32_t SBOZ = (C / 8) * 128;              // W_in operand: [rows x C], LBO 128
  static constexpr uint32_t WIN = 2 * 4 * SBOZ;                // W_in slice: 2 terms x [32 x C]
  static constexpr uint32_t WOUT_T = (C / 8) * 256;            // one term of the [C x 16] W_out slice
  static constexpr uint32_t WO = GF_DW_BYTES + 2 * WOUT_T;     // dw taps + W_out slice (contiguous in the blob)
  static constexpr uint32_t SLICE = WIN + WO;
  // shared memory carve-up
  // (the LN(x) operand of the tile lives in TENSOR memory: columns [0, 2C) = {tile 0 hi, lo, tile 1 hi, lo})
  static constexpr uint32_t OFF_U = 0;
  static constexpr uint32_t OFF_G = OFF_U + GF_U_BUFS * GF_U_BYTES;
  static constexpr uint32_t OFF_WIN = OFF_G + 2 * 2 * GF_G_TILE;
  static constexpr uint32_t OFF_WO = OFF_WIN + GF_W_SLOTS * WIN;
  static constexpr uint32_t OFF_GB = OFF_WO + GF_W_SLOTS * WO; // gamma, beta
  static constexpr uint32_t OFF_ST = OFF_GB + 2 * C * 4;       // partial statistics [4][128][2]
  static constexpr uint32_t TOTAL = OFF_ST + 4 * 128 * 2 * 4;
};

// Per-warp wait / section cycle counters of CTA 0 (measurement only: kernel instantiation PROF, p.debug & 16).
__device__ unsigned long long gf_prof[GF_WARPS * 8];

template <int C, bool PROF>
__global__ void __launch_bounds__(GF_THREADS, 1) gdfn_fwd_kernel(const rcot_gdfn_params p, const int tiles_x,
                                                                const int tiles_per_img, const int total_tiles) {
  using L = GfLayout<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  // TMA completions: winbar, wobar.  tcgen05.commit: ubar (GEMM-1 done), gbar (GEMM-2 done: g slot + W_out slot free).
  // drain -> issuer: dbar (U TMEM buffer in registers).  drain -> stencil: ufull.  stencil -> drain: uempty.
  // stencil -> issuer: zbar (Z of the tile in TMEM), sbar (g slice written), ybar (Y drained by the epilogue).
  __shared__ uint64_t winbar[GF_W_SLOTS], wobar[GF_W_SLOTS], ubar[2], gbar[2], dbar[2], sbar[2];
  __shared__ uint64_t ufull[GF_U_BUFS], uempty[GF_U_BUFS], zbar, ybar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, W = p.W, HWp = H * W, hid = p.hid;
  const int NS = (hid + GF_HS - 1) / GF_HS;
  const bool LN = p.ln_stats != nullptr;
  float* gb = reinterpret_cast<float*>(smem + L::OFF_GB);
  float* stp = reinterpret_cast<float*>(smem + L::OFF_ST);

  if (LN)
    for (int c = tid; c < C; c += GF_THREADS) {
      gb[c] = __ldg(p.ln_gamma + c);
      gb[C + c] = __ldg(p.ln_beta + c);
    }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < GF_W_SLOTS; ++i) {
      mbar_init(&winbar[i], 1);
      mbar_init(&wobar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ubar[i], 1);
      mbar_init(&gbar[i], 1);
      mbar_init(&dbar[i], GF_D_WARPS);
      mbar_init(&sbar[i], GF_S_WARPS / 2);
    }
    for (int i = 0; i < GF_U_BUFS; ++i) {
      mbar_init(&ufull[i], GF_D_WARPS);
      mbar_init(&uempty[i], GF_S_WARPS / 2);
    }
    mbar_init(&zbar, GF_S_WARPS);
    mbar_init(&ybar, GF_S_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = PROF ? clock64() : 0;
  auto tw = [&](int site, uint64_t* bar, uint32_t ph) {       // timed mbarrier wait
    if (PROF) {
      const long long t0 = clock64();
      mbar_wait(bar, ph);
      acc[site] += clock64() - t0;
    } else {
      mbar_wait(bar, ph);
    }
  };
  // TMEM columns: Z (A operand of GEMM-1) [0, 2C), U buffers [2C, 2C + 128), Y [2C + 128, 3C + 128)   (<= 416 of 512)
  const uint32_t tmem_u = tmem + 2 * C;
  const uint32_t tmem_y = tmem_u + 128;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_slices = my_tiles * NS;

  if (warp == GF_ISSUER) {
    // ================================================================ GEMM-1 issuer: W_in ring (TMA) + U_n = Z . W_in[n]^T
    // Never waits for the stencil: it runs as far ahead as the two TMEM U buffers (dbar) and Z (zbar) allow.
    const uint8_t* blob = reinterpret_cast<const uint8_t*>(p.wblob);
    const uint32_t idesc1 = make_idesc_bf16(128, 32);
    const uint32_t win_base = smem_u32(smem + L::OFF_WIN);
    const uint64_t dz_t = make_sdesc(0, 128, L::SBOZ);            // everything but the 14-bit start address
    auto load_win = [&](int n, int s) {
      const int ws = n & (GF_W_SLOTS - 1);
      mbar_arrive_expect_tx(&winbar[ws], L::WIN);
      bulk_g2s(smem + L::OFF_WIN + ws * L::WIN, blob + (size_t)s * L::SLICE, L::WIN, &winbar[ws]);
    };
    if (lane == 0) {
      if (total_slices > 0) load_win(0, 0);
      if (total_slices > 1) load_win(1, 1 % NS);
    }
    __syncwarp();
    int n = 0, s2 = 2 % NS;                                         // s2 = slice-in-tile of global slice n + 2
    for (int ti = 0; ti < my_tiles; ++ti) {
      tw(0, &zbar, (uint32_t)ti & 1);                               // Z of this tile is in tensor memory
      for (int s = 0; s < NS; ++s, ++n) {
        tw(1, &winbar[n & (GF_W_SLOTS - 1)], (uint32_t)(n >> 2) & 1);
        if (n >= 2) tw(2, &dbar[n & 1], (uint32_t)((n - 2) >> 1) & 1);   // TMEM buffer read out; GEMM-1(n-2) retired
        // W_in of slice n+2 into the slot of slice n-2
        if (n + 2 < total_slices && lane == 0) load_win(n + 2, s2);
        s2 = (s2 + 1 == NS) ? 0 : s2 + 1;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t wb = (win_base + (uint32_t)(n & (GF_W_SLOTS - 1)) * L::WIN) >> 4;
          if (!(p.debug & 1)) {
            // the two 128-row tiles have independent accumulators: alternating them halves the exposed
            // accumulate-after-accumulate latency of these short (N = 32, K = 16) MMAs
            const uint32_t d0 = tmem_u + (uint32_t)(n & 1) * 64, d1 = d0 + 32;
            const uint32_t ah0 = tmem, al0 = ah0 + C / 2, ah1 = tmem + C, al1 = ah1 + C / 2;   // A in TMEM: 8 columns per k16
#pragma unroll
            for (int ks = 0; ks < C / 16; ++ks) {
              const uint64_t dbh = dz_t | (uint64_t)((wb + ks * 16) & 0x3FFFu);
              const uint64_t dbl = dz_t | (uint64_t)((wb + (4 * L::SBOZ >> 4) + ks * 16) & 0x3FFFu);
              tc_mma_bf16_ts(d0, ah0 + ks * 8, dbh, idesc1, ks == 0 ? 0u : 1u);
              tc_mma_bf16_ts(d1, ah1 + ks * 8, dbh, idesc1, ks == 0 ? 0u : 1u);
              tc_mma_bf16_ts(d0, al0 + ks * 8, dbh, idesc1, 1u);
              tc_mma_bf16_ts(d1, al1 + ks * 8, dbh, idesc1, 1u);
              tc_mma_bf16_ts(d0, ah0 + ks * 8, dbl, idesc1, 1u);
              tc_mma_bf16_ts(d1, ah1 + ks * 8, dbl, idesc1, 1u);
            }
          }
          tc_commit(&ubar[n & 1]);
        }
        __syncwarp();
      }
    }
  } else if (warp == GF_ISSUER + 1) {
    // ================================================================ GEMM-2 issuer: W_out / tap ring + Y += g_n . W_out[:, n]^T
    const uint8_t* blob = reinterpret_cast<const uint8_t*>(p.wblob);
    const uint32_t idesc2 = make_idesc_bf16(128, C);
    const uint32_t g_base = smem_u32(smem + L::OFF_G), wo_base = smem_u32(smem + L::OFF_WO);
    const uint64_t dg_t = make_sdesc(0, 128, GF_G_SBO), dw_t = make_sdesc(0, 128, 256);
    auto load_wo = [&](int n, int s) {
      const int ws = n & (GF_W_SLOTS - 1);
      mbar_arrive_expect_tx(&wobar[ws], L::WO);
      bulk_g2s(smem + L::OFF_WO + ws * L::WO, blob + (size_t)s * L::SLICE + L::WIN, L::WO, &wobar[ws]);
    };
    if (lane == 0)
      for (int i = 0; i < GF_W_SLOTS && i < total_slices; ++i) load_wo(i, i % NS);
    __syncwarp();
    int m = 0, s3 = 3 % NS;                                         // s3 = slice-in-tile of global slice m + 3
    for (int ti = 0; ti < my_tiles; ++ti)
      for (int s = 0; s < NS; ++s, ++m) {
        tw(3, &sbar[m & 1], (uint32_t)(m >> 1) & 1);                // g(m) written
        if (s == 0 && ti > 0) tw(4, &ybar, (uint32_t)(ti - 1) & 1); // the first GEMM-2 of a tile overwrites Y
        tc_fence_after();
        if (elect_one()) {
          const uint32_t gh = (g_base + (uint32_t)(m & 1) * 2 * GF_G_TILE) >> 4, gl = gh + (GF_G_TILE >> 4);
          const uint32_t wh = (wo_base + (uint32_t)(m & (GF_W_SLOTS - 1)) * L::WO + GF_DW_BYTES) >> 4, wl = wh + (L::WOUT_T >> 4);
          if (!(p.debug & 1)) {
            const uint64_t dgh = dg_t | (uint64_t)(gh & 0x3FFFu), dgl = dg_t | (uint64_t)(gl & 0x3FFFu);
            const uint64_t dwh = dw_t | (uint64_t)(wh & 0x3FFFu), dwl = dw_t | (uint64_t)(wl & 0x3FFFu);
            tc_mma_bf16(tmem_y, dgh, dwh, idesc2, s == 0 ? 0u : 1u);
            tc_mma_bf16(tmem_y, dgl, dwh, idesc2, 1u);
            tc_mma_bf16(tmem_y, dgh, dwl, idesc2, 1u);
          }
          tc_commit(&gbar[m & 1]);
        }
        __syncwarp();
        // W_out + taps of slice m+3 into the slot of slice m-1 (its stencil is done: sbar(m-1); GEMM-2(m-1) retired)
        if (m >= 1 && m + GF_W_SLOTS - 1 < total_slices) {
          tw(5, &gbar[(m - 1) & 1], (uint32_t)((m - 1) >> 1) & 1);
          if (lane == 0) load_wo(m + GF_W_SLOTS - 1, s3);
          __syncwarp();
        }
        s3 = (s3 + 1 == NS) ? 0 : s3 + 1;
      }
  } else if (warp >= GF_S_WARPS) {
    // ================================================================ 4 drain warps: U_n  TMEM -> registers -> shared
    const int q = warp & 3;                                        // TMEM lane quarter (GF_S_WARPS % 4 == 0)
    const int hp0 = q * 32 + lane, hp1 = 128 + hp0;                // halo pixels of this thread in MMA tiles 0 and 1
    const int hy0 = hp0 / GF_HW, hx0 = hp0 - hy0 * GF_HW;
    const int hy1 = hp1 / GF_HW, hx1 = hp1 - hy1 * GF_HW;
    const bool two = q < 2;                                        // warp-uniform: tile 1 holds halo pixels 128..179 only
    int n = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int b = t / tiles_per_img, tr = t - b * tiles_per_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const int y0 = ty * GF_TH, x0 = tx * GF_TW;
      if (t + (int)gridDim.x < total_tiles) {
        // pull the x rows of the NEXT tile's halo region towards L2 (its Z is built ~a tile time from now)
        const int t2 = t + gridDim.x;
        const int b2 = t2 / tiles_per_img, tr2 = t2 - b2 * tiles_per_img;
        const int ty2 = tr2 / tiles_x, tx2 = tr2 - ty2 * tiles_x;
        const int gx0 = max(tx2 * GF_TW - 1, 0);
        const float* xn = p.x + (size_t)b2 * p.x_bs + gx0;
        for (int idx = (warp - GF_S_WARPS) * 32 + lane; idx < C * GF_HH; idx += GF_D_WARPS * 32) {
          const int ch = idx / GF_HH, gy = ty2 * GF_TH - 1 + (idx - ch * GF_HH);
          if ((unsigned)gy < (unsigned)H) {
            const float* a0 = xn + (size_t)ch * HWp + (size_t)gy * W;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + 16));
          }
        }
      }
      for (int s = 0; s < NS; ++s, ++n) {
        tw(0, &ubar[n & 1], (uint32_t)(n >> 1) & 1);              // GEMM-1(n) retired
        tc_fence_after();
        const uint32_t ta = tmem_u + ((uint32_t)(q * 32) << 16) + (uint32_t)(n & 1) * 64;
        const int ub = n % GF_U_BUFS;
        float* Ub = reinterpret_cast<float*>(smem + L::OFF_U + ub * GF_U_BYTES);
        auto put = [&](const uint32_t (&r)[32], int hp, int hy, int hx) {
          if (hp >= GF_NHP || (p.debug & 4)) return;
          float* up = Ub + hy * GF_RS + hx;
#pragma unroll
          for (int i = 0; i < 32; ++i) up[i * GF_CS] = __uint_as_float(r[i]);
        };
        uint32_t r[32];
        tmem_ld32_nowait(ta, r);
        tmem_ld_wait();
        if (n >= GF_U_BUFS) tw(1, &uempty[ub], (uint32_t)(n / GF_U_BUFS - 1) & 1);   // stencil(n-3) done with it
        put(r, hp0, hy0, hx0);
        if (two) {
          tmem_ld32_nowait(ta + 32, r);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dbar[n & 1]);                  // the TMEM buffer may be overwritten by GEMM-1(n+2)
        if (two) put(r, hp1, hy1, hx1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&ufull[ub]);
      }
    }
  } else {
    // ================================================================ 16 stencil warps (two groups of 8)
    // Z phase of one tile: LN(x) of the halo tile as the bf16 hi/lo operand; arrives on zbar
    auto produce_z = [&](int t) {
      const int b = t / tiles_per_img, tr = t - b * tiles_per_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const int y0 = ty * GF_TH, x0 = tx * GF_TW;
      const int hp = tid & 255, half = tid >> 8;
      const int hy = hp / GF_HW, hx = hp - hy * GF_HW;
      const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
      const bool inimg = hp < GF_NHP && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
      const int mt = hp >> 7;                          // warp-uniform: hp = 32 * (warp & 7) + lane
      // this warp's TMEM lanes are 32 * (warp & 3) .. +31 = rows (hp & 127) of tile mt; columns of channel group kg:
      //   hi: mt*C + 4*kg .. +3,   lo: mt*C + C/2 + 4*kg .. +3      (8 channels = 4 packed columns)
      const uint32_t zaddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * C);
      constexpr int NG = C / 16;                      // 8-channel groups per thread (half of the channels)
      float v[NG][8];
      if (inimg) {
        const float* xp = p.x + (size_t)b * p.x_bs + (size_t)gy * W + gx + (size_t)(half * (C / 2)) * HWp;
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) v[g][i] = __ldg(xp + (size_t)(g * 8 + i) * HWp);
        if (LN) {
          const float2 st = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)b * HWp + gy * W + gx);
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const float* gp = gb + half * (C / 2) + g * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[g][i] = (v[g][i] - st.x) * st.y * gp[i] + gp[C + i];
          }
        }
      } else {                                         // outside the image (or beyond the 180 halo pixels): zero rows
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) v[g][i] = 0.f;
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {                  // every lane takes part in the (warp-collective) TMEM stores
        uint4 hi, lo;
        split8(v[g], hi, lo);
        const uint32_t kc = (uint32_t)(half * (C / 16) + g) * 4;
        tmem_st4(zaddr + kc, hi.x, hi.y, hi.z, hi.w);
        tmem_st4(zaddr + C / 2 + kc, lo.x, lo.y, lo.z, lo.w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&zbar);
    };

    const int grp = warp >> 3, wg = warp & 7;
    // stencil mapping: warp = (output row pair rp, pair half); lane = (pair jj of the half, 4-pixel strip xq): a
    // quarter-warp reads 8 different channels at one strip (conflict-free 16-byte accesses, GF_CS) and the 8 pairs of
    // a lane group are the 8 contiguous k values of one operand row
    const int rp = wg >> 1, j = (wg & 1) * 8 + (lane & 7), xq = lane >> 3;
    int n = 0, ti = 0;
    if (blockIdx.x < total_tiles) produce_z(blockIdx.x);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const int b = t / tiles_per_img, tr = t - b * tiles_per_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const int y0 = ty * GF_TH, x0 = tx * GF_TW;
      const float* xb = p.x + (size_t)b * p.x_bs;
      const int nl = n + NS - 1;                                     // last slice of this tile
      for (int s = 0; s < NS; ++s, ++n) {
        if ((n & 1) != grp) continue;
        if (s >= NS - 2 && t + (int)gridDim.x < total_tiles) {
          // This group's last slice of the tile.  Once GEMM-1(nl) has retired (ufull(nl): a peek for the group that
          // does not own slice nl -- it has seen ufull(nl-1), and the drain warps fill in order, so the parity cannot
          // alias) nothing reads Z any more: build the NEXT tile's Z now, so that its first GEMM-1s and drains run
          // under the last stencils and the epilogue instead of after them.
          const long long tz1 = PROF ? clock64() : 0;
          tw(3, &ufull[nl % GF_U_BUFS], (uint32_t)(nl / GF_U_BUFS) & 1);
          tc_fence_after();
          produce_z(t + gridDim.x);
          if (PROF) acc[4] += clock64() - tz1;
        }
        const int ub = n % GF_U_BUFS;
        tw(0, &ufull[ub], (uint32_t)(n / GF_U_BUFS) & 1);                          // U_n is in shared memory
        tw(1, &wobar[n & (GF_W_SLOTS - 1)], (uint32_t)(n / GF_W_SLOTS) & 1);      // depthwise taps of slice n
        if (n >= 2) tw(2, &gbar[n & 1], (uint32_t)((n - 2) >> 1) & 1);            // GEMM-2(n-2) retired: g slot free
        if (!(p.debug & 2)) {
          const float* wdw = reinterpret_cast<const float*>(smem + L::OFF_WO + (n & (GF_W_SLOTS - 1)) * L::WO);
          float wv[20];                    // taps of the pair: a0..a8, b0..b8 (+2 pad) as five 16-byte loads
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float4 t4 = lds128(wdw + j * 20 + 4 * i);
            wv[4 * i] = t4.x; wv[4 * i + 1] = t4.y; wv[4 * i + 2] = t4.z; wv[4 * i + 3] = t4.w;
          }
          const float* wa = wv;
          const float* wb = wv + 9;
          float a[2][4], bb[2][4];
#pragma unroll
          for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int i = 0; i < 4; ++i) a[o][i] = bb[o][i] = 0.f;
          const float* ua = reinterpret_cast<const float*>(smem + L::OFF_U + ub * GF_U_BYTES) + j * GF_CS +
                            (2 * rp) * GF_RS + 4 * xq;
          const float* ubp = ua + 16 * GF_CS;
#pragma unroll
          for (int hr = 0; hr < 4; ++hr) {
            // columns 4xq .. 4xq+7 of halo row 2rp + hr (6 are used; the row stride of 20 keeps the second access in
            // the row); the row feeds output row 0 with dy = hr and output row 1 with dy = hr - 1
            const float4 a4 = lds128(ua + hr * GF_RS), a2 = lds128(ua + hr * GF_RS + 4);
            const float4 b4 = lds128(ubp + hr * GF_RS), b2 = lds128(ubp + hr * GF_RS + 4);
            const float va[6] = {a4.x, a4.y, a4.z, a4.w, a2.x, a2.y};
            const float vb[6] = {b4.x, b4.y, b4.z, b4.w, b2.x, b2.y};
            if (p.save_u && (hr == 1 || hr == 2) && s * GF_HS + j < hid) {
              // halo row 2rp + hr = core row 2rp + hr - 1; halo columns 4xq+1 .. 4xq+4 = core columns 4xq .. 4xq+3
              float* su = p.save_u + (size_t)b * p.u_bs + (size_t)(s * GF_HS + j) * HWp +
                          (size_t)(y0 + 2 * rp + hr - 1) * W + x0 + 4 * xq;
              *reinterpret_cast<float4*>(su) = make_float4(va[1], va[2], va[3], va[4]);
              *reinterpret_cast<float4*>(su + (size_t)hid * HWp) = make_float4(vb[1], vb[2], vb[3], vb[4]);
            }
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              const int dy = hr - o;
              if (dy >= 0 && dy <= 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                  for (int dx = 0; dx < 3; ++dx) {
                    a[o][i] = fmaf(wa[dy * 3 + dx], va[i + dx], a[o][i]);
                    bb[o][i] = fmaf(wb[dy * 3 + dx], vb[i + dx], bb[o][i]);
                  }
              }
            }
          }
          uint8_t* gh = smem + L::OFF_G + (n & 1) * 2 * GF_G_TILE + (j >> 3) * 128 + (j & 7) * 2;
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            float g[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] = gelu_fast(a[o][i]) * bb[o][i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int pr = (2 * rp + o) * GF_TW + 4 * xq + i;
              const __nv_bfloat16 hi = __float2bfloat16_rn(g[i]);
              const __nv_bfloat16 lo = __float2bfloat16_rn(g[i] - __bfloat162float(hi));
              const uint32_t off = (uint32_t)(pr >> 3) * GF_G_SBO + (uint32_t)(pr & 7) * 16;
              *reinterpret_cast<__nv_bfloat16*>(gh + off) = hi;
              *reinterpret_cast<__nv_bfloat16*>(gh + GF_G_TILE + off) = lo;
            }
            if (p.save_g && s * GF_HS + j < hid) {
              float* sg = p.save_g + (size_t)b * p.g_bs + (size_t)(s * GF_HS + j) * HWp + (size_t)(y0 + 2 * rp + o) * W +
                          x0 + 4 * xq;
              *reinterpret_cast<float4*>(sg) = make_float4(g[0], g[1], g[2], g[3]);
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sbar[n & 1]);
          mbar_arrive(&uempty[ub]);
        }
      }
      // ---- both groups meet (after it, the non-owner of slice nl may peek at gbar(nl): the owner waited for
      //      GEMM-2(nl-2) before it wrote g(nl), so the parity cannot alias)
      worker_sync();
      // ---- epilogue: Y (+ x) -> HBM, LayerNorm statistics of y
      const long long te0 = PROF ? clock64() : 0;
      {
        const int q = warp & 3, cg = warp >> 2;
        const int pr = q * 32 + lane, r = pr >> 4, cx = pr & 15;
        const size_t pix = (size_t)(y0 + r) * W + x0 + cx;
        const float* xr = xb + pix;
        float* yo = p.y + (size_t)b * p.y_bs + pix;
        const bool want_stats = p.stats_out != nullptr;
        const float shift = want_stats ? __ldg(xr) : 0.f;
        float s1 = 0.f, s2 = 0.f;
        constexpr int NGI = (C / 8 + 3) / 4;                       // 8-channel groups of this warp: gi = cg + 4 * k < C / 8
        float res[NGI][8];
#pragma unroll
        for (int k = 0; k < NGI; ++k)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            res[k][i] = (p.residual && cg + 4 * k < C / 8) ? __ldg(xr + (size_t)((cg + 4 * k) * 8 + i) * HWp) : 0.f;
        tw(5, &gbar[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1);   // the last GEMM-2 of the tile has retired
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < NGI; ++k) {
          const int gi = cg + 4 * k;
          if (gi >= C / 8) break;                                  // warp-uniform
          uint32_t rr[8];
          tmem_ld8_nowait(tmem_y + ((uint32_t)(q * 32) << 16) + gi * 8, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float yv = __uint_as_float(rr[i]) + res[k][i];
            yo[(size_t)(gi * 8 + i) * HWp] = yv;
            const float d = yv - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ybar);                         // Y drained: the next tile's GEMM-2 may overwrite it
        if (want_stats) {
          stp[(cg * 128 + pr) * 2] = s1;
          stp[(cg * 128 + pr) * 2 + 1] = s2;
          worker_sync();
          if (cg == 0) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              t1 += stp[(k * 128 + pr) * 2];
              t2 += stp[(k * 128 + pr) * 2 + 1];
            }
            const float inv = 1.f / (float)C;
            const float m = t1 * inv;
            const float var = fmaxf(t2 * inv - m * m, 0.f);
            reinterpret_cast<float2*>(p.stats_out)[(size_t)b * HWp + pix] = make_float2(shift + m, 1.0f / sqrtf(var + 1e-5f));
          }
          // (stp is rewritten by the next tile's epilogue only after the worker barrier that follows its slices)
        }
      }
      if (PROF) acc[7] += clock64() - te0;
    }
  }
  if (PROF && blockIdx.x == 0 && lane == 0) {
    acc[6] = clock64() - t_begin;
    for (int i = 0; i < 8; ++i) gf_prof[warp * 8 + i] = (unsigned long long)acc[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------- weight blob
template <int C>
__global__ void gdfn_pack_kernel(const float* __restrict__ w_in, const float* __restrict__ w_dw,
                                 const float* __restrict__ w_out, uint8_t* __restrict__ blob, int hid) {
  using L = GfLayout<C>;
  const int s = blockIdx.x;
  uint8_t* dst = blob + (size_t)s * L::SLICE;
  // W_in slice: rows 0..15 = a channels s*16+i, rows 16..31 = b channels hid + s*16 + i
  for (int e = threadIdx.x; e < 32 * C; e += blockDim.x) {
    const int i = e / C, k = e - i * C;
    const int pair = s * GF_HS + (i & 15);
    float w = 0.f;
    if (pair < hid) w = w_in[(size_t)((i < 16) ? pair : hid + pair) * C + k];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const uint32_t off = (uint32_t)(i >> 3) * L::SBOZ + (uint32_t)(k >> 3) * 128 + (uint32_t)(i & 7) * 16 + (uint32_t)(k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(dst + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(dst + 4 * L::SBOZ + off) = lo;
  }
  float* dw = reinterpret_cast<float*>(dst + L::WIN);
  for (int e = threadIdx.x; e < 16 * 20; e += blockDim.x) {
    const int jp = e / 20, tp = e - jp * 20;           // pair of the slice, slot: 0..8 a-taps, 9..17 b-taps, 18..19 pad
    const int pair = s * GF_HS + jp;
    float w = 0.f;
    if (pair < hid && tp < 18) w = w_dw[(size_t)(tp < 9 ? pair : hid + pair) * 9 + (tp < 9 ? tp : tp - 9)];
    dw[e] = w;
  }
  uint8_t* wo = dst + L::WIN + GF_DW_BYTES;
  for (int e = threadIdx.x; e < C * 16; e += blockDim.x) {
    const int nrow = e >> 4, kk = e & 15;
    const int pair = s * GF_HS + kk;
    const float w = (pair < hid) ? w_out[(size_t)nrow * hid + pair] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const uint32_t off = (uint32_t)(nrow >> 3) * 256 + (uint32_t)(kk >> 3) * 128 + (uint32_t)(nrow & 7) * 16 + (uint32_t)(kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(wo + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(wo + L::WOUT_T + off) = lo;
  }
}

static int gf_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int C, bool PROF>
static int launch_gdfn(const rcot_gdfn_params& p, cudaStream_t stream) {
  using L = GfLayout<C>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gdfn_fwd_kernel<C, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
    if (e != cudaSuccess) {
      set_error("gdfn_fwd: cudaFuncSetAttribute(%u bytes): %s", L::TOTAL, cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int tiles_x = p.W / GF_TW, tiles_y = p.H / GF_TH;
  const int tpi = tiles_x * tiles_y;
  const long total = (long)tpi * p.B;
  const int grid = (int)(total < gf_num_sms() ? total : gf_num_sms());
  gdfn_fwd_kernel<C, PROF><<<grid, GF_THREADS, L::TOTAL, stream>>>(p, tiles_x, tpi, (int)total);
  return check_launch("gdfn_fwd");
}

}  // namespace rcot

using namespace rcot;

extern "C" int rcot_gdfn_supported(int C, int H, int W) {
  return (C == 48 || C == 96) && H % GF_TH == 0 && W % GF_TW == 0 && H > 0 && W > 0;
}

extern "C" size_t rcot_gdfn_blob_bytes(int C, int hid) {
  const size_t ns = (size_t)(hid + GF_HS - 1) / GF_HS;
  if (C == 48) return ns * GfLayout<48>::SLICE;
  if (C == 96) return ns * GfLayout<96>::SLICE;
  return 0;
}

extern "C" int rcot_gdfn_pack(const float* w_in, const float* w_dw, const float* w_out, void* blob, int C, int hid,
                              rcot_stream_t st) {
  RCOT_REQUIRE(w_in && w_dw && w_out && blob && hid > 0, "gdfn_pack: bad arguments");
  RCOT_REQUIRE(C == 48 || C == 96, "gdfn_pack: the fused GDFN kernel is built for C = 48 and 96 (got %d)", C);
  const int ns = (hid + GF_HS - 1) / GF_HS;
  if (C == 48)
    gdfn_pack_kernel<48><<<ns, 256, 0, (cudaStream_t)st>>>(w_in, w_dw, w_out, reinterpret_cast<uint8_t*>(blob), hid);
  else
    gdfn_pack_kernel<96><<<ns, 256, 0, (cudaStream_t)st>>>(w_in, w_dw, w_out, reinterpret_cast<uint8_t*>(blob), hid);
  return check_launch("gdfn_pack");
}

extern "C" int rcot_gdfn_fwd(const rcot_gdfn_params* pp, rcot_stream_t st) {
  RCOT_REQUIRE(pp != nullptr, "gdfn_fwd: null params");
  const rcot_gdfn_params& p = *pp;
  RCOT_REQUIRE(p.x && p.y && p.wblob, "gdfn_fwd: null tensor pointer");
  RCOT_REQUIRE(p.B > 0 && p.hid > GF_HS, "gdfn_fwd: bad sizes (hid must exceed one slice of %d pairs)", GF_HS);
  RCOT_REQUIRE(rcot_gdfn_supported(p.C, p.H, p.W), "gdfn_fwd: needs C in {48, 96}, H %% 8 == 0, W %% 16 == 0 (got C=%d %dx%d)",
               p.C, p.H, p.W);
  if (p.ln_stats) RCOT_REQUIRE(p.ln_gamma && p.ln_beta, "gdfn_fwd: LayerNorm needs gamma and beta");
  RCOT_REQUIRE((long)p.B * (p.H / GF_TH) * (p.W / GF_TW) < (1L << 31), "gdfn_fwd: too many tiles");
  if (p.save_g) RCOT_REQUIRE((reinterpret_cast<uintptr_t>(p.save_g) & 15) == 0 && p.g_bs % 4 == 0, "gdfn_fwd: save_g alignment");
  if (p.save_u) RCOT_REQUIRE((reinterpret_cast<uintptr_t>(p.save_u) & 15) == 0 && p.u_bs % 4 == 0, "gdfn_fwd: save_u alignment");
  if (p.debug & 16) return p.C == 48 ? launch_gdfn<48, true>(p, (cudaStream_t)st) : launch_gdfn<96, true>(p, (cudaStream_t)st);
  return p.C == 48 ? launch_gdfn<48, false>(p, (cudaStream_t)st) : launch_gdfn<96, false>(p, (cudaStream_t)st);
}

/* Measurement aid: cycle counters of CTA 0 of the last rcot_gdfn_fwd launched with debug & 16 ([22 warps][8] uint64;
 * per warp: wait sites 0..5, 6 = whole kernel, 7 = epilogue body).  Synchronises the device. */
extern "C" int rcot_gdfn_profile_read(unsigned long long* out, int n) {
  RCOT_REQUIRE(out && n == GF_WARPS * 8, "gdfn_profile_read: expects %d counters", GF_WARPS * 8);
  cudaError_t e = cudaMemcpyFromSymbol(out, gf_prof, sizeof(unsigned long long) * n);
  if (e != cudaSuccess) {
    set_error("gdfn_profile_read: %s", cudaGetErrorString(e));
    return RCOT_ERR_CUDA;
  }
  return RCOT_OK;
}
