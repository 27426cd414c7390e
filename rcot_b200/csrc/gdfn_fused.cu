// gdfn_fused.cu -- GDFN forward as ONE kernel (Net_Restormer.py:80-85 with the block's LayerNorm and residual):
//     y = x + W_out . ( gelu(dw3x3(W_in . LN(x))[:hid]) * dw3x3(W_in . LN(x))[hid:] )
// The 5.3 C-wide hidden tensor u never reaches HBM: a CTA owns an 8 x 16-pixel tile of one image (+1-pixel halo = 180
// pixels = two 128-row MMA tiles) and walks the hidden dimension in SLICES of 16 (a, b) channel pairs:
//   Z   : LN(x) of the halo tile as the bf16 hi/lo A operand [256 x C], built once per tile (halo pixels outside the
//         image become zero rows, so their u is the zero padding the depthwise conv expects -- not LN(0) = beta);
//   per slice s:  GEMM-1  U_s[256 x 32] = Z . W_in[slice]^T          tcgen05 -> TMEM (double buffered)
//                 drain   TMEM -> shared memory as [channel][10][18] fp32
//                 stencil a = dw(U_s[:16]), b = dw(U_s[16:]), g = gelu(a) * b  on the 128 core pixels (CUDA cores)
//                         -> bf16 hi/lo A operand [128 x 16]
//                 GEMM-2  Y[128 x C] += g . W_out[:, slice]^T         tcgen05, accumulates over the slices in TMEM
//   epilogue: Y (+ x) -> HBM, plus the per-pixel LayerNorm statistics the next block's LN1 needs.
// One thread issues the MMAs asynchronously: GEMM-1 of slice s+1 and GEMM-2 of slice s-1 run on the tensor pipe while
// all 16 warps drain / convolve slice s, so the phases need only two block barriers per slice.  Weight slices
// (pre-packed by rcot_gdfn_pack: operand images + the 32 x 9 depthwise taps) stream through the TMA engine
// (cp.async.bulk) two slices ahead.  bf16x3 split products (hi*hi + lo*hi + hi*lo) as everywhere else: fp32-class.
// Algorithmic HBM bytes: (1 + 180/128 halo re-read, mostly L2 hits) C + C per pixel instead of ~19 C for the three
// unfused launches.  Optional outputs u / g keep the existing (unfused) backward fed when it wants them saved.
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "gelu.cuh"
#include "tc.cuh"

namespace rcot {

constexpr int GF_TH = 8, GF_TW = 16;                 // core tile (pixels)
constexpr int GF_HH = GF_TH + 2, GF_HW = GF_TW + 2;  // halo tile
constexpr int GF_NHP = GF_HH * GF_HW;                // 180 halo pixels
constexpr int GF_RS = 20;                            // shared-memory row stride of a halo row (floats)
constexpr int GF_CS = 228;                           // channel stride (floats) >= 10 * 20 and == 4 (mod 32): the 8 pairs a
                                                     // quarter-warp reads in one 16-byte access fall into 8 distinct bank groups
constexpr int GF_HS = 16;                            // (a, b) pairs per hidden slice
constexpr int GF_WORKER_WARPS = 16;                  // drain / stencil / epilogue warps
constexpr int GF_THREADS = (GF_WORKER_WARPS + 1) * 32;   // + the issuer warp (TMA weight ring, every tcgen05.mma)
constexpr int GF_WIN_SLOTS = 3;                      // W_in slice ring
constexpr uint32_t GF_G_SBO = 272;                   // g operand: 8-row group stride (256 + 16 B pad against bank conflicts)
constexpr uint32_t GF_G_TILE = 16 * GF_G_SBO;        // one term of the [128 x 16] g operand
constexpr uint32_t GF_DW_BYTES = 16 * 20 * sizeof(float);   // per pair: 9 a-taps, 9 b-taps, 2 pad (16-byte loads)

template <int C>
struct GfLayout {
  static constexpr uint32_t SBOZ = (C / 8) * 128;              // Z / W_in operands: [rows x C], LBO 128
  static constexpr uint32_t ZT = 16 * SBOZ;                    // one term of one 128-row Z tile
  static constexpr uint32_t WIN = 2 * 4 * SBOZ;                // W_in slice: 2 terms x [32 x C]
  static constexpr uint32_t WOUT_T = (C / 8) * 256;            // one term of the [C x 16] W_out slice
  static constexpr uint32_t WO = GF_DW_BYTES + 2 * WOUT_T;     // dw taps + W_out slice (contiguous in the blob)
  static constexpr uint32_t SLICE = WIN + WO;
  // shared memory carve-up
  // (the LN(x) operand of the tile lives in TENSOR memory: columns [0, 2C) = {tile 0 hi, lo, tile 1 hi, lo})
  static constexpr uint32_t OFF_U = 0;
  static constexpr uint32_t OFF_G = OFF_U + 32 * GF_CS * 4;
  static constexpr uint32_t OFF_WIN = OFF_G + 2 * 2 * GF_G_TILE;
  static constexpr uint32_t OFF_WO = OFF_WIN + GF_WIN_SLOTS * WIN;
  static constexpr uint32_t OFF_GB = OFF_WO + 4 * WO;          // gamma, beta
  static constexpr uint32_t OFF_ST = OFF_GB + 2 * C * 4;       // partial statistics [4][128][2]
  static constexpr uint32_t TOTAL = OFF_ST + 4 * 128 * 2 * 4;
};

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// 16-byte shared-memory load the compiler may not narrow (a quarter-warp phase of 8 lanes is conflict-free here; the
// 8-byte form's half-warp phase is not).
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

// Named barrier among the 16 worker warps only (the issuer warp never joins it).
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <int C>
__global__ void __launch_bounds__(GF_THREADS, 1) gdfn_fwd_kernel(const rcot_gdfn_params p, const int tiles_x,
                                                                const int tiles_per_img, const int total_tiles) {
  using L = GfLayout<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  // TMA completions: winbar, wobar.  tcgen05.commit: ubar (U slice ready), gbar (GEMM-2 done: g slot + W_out slot free).
  // workers -> issuer (one arrival per worker warp): zbar (Z of the tile ready), dbar (U TMEM buffer drained),
  // sbar (g slice written), ybar (Y drained by the epilogue).
  __shared__ uint64_t winbar[GF_WIN_SLOTS], wobar[4], ubar[2], gbar[2], zbar, dbar[2], sbar[2], ybar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, W = p.W, HWp = H * W, hid = p.hid;
  const int NS = (hid + GF_HS - 1) / GF_HS;
  const bool LN = p.ln_stats != nullptr;
  float* Usm = reinterpret_cast<float*>(smem + L::OFF_U);
  float* gb = reinterpret_cast<float*>(smem + L::OFF_GB);
  float* stp = reinterpret_cast<float*>(smem + L::OFF_ST);

  if (LN)
    for (int c = tid; c < C; c += GF_THREADS) {
      gb[c] = __ldg(p.ln_gamma + c);
      gb[C + c] = __ldg(p.ln_beta + c);
    }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < GF_WIN_SLOTS; ++i) mbar_init(&winbar[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&wobar[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ubar[i], 1);
      mbar_init(&gbar[i], 1);
      mbar_init(&dbar[i], GF_WORKER_WARPS);
      mbar_init(&sbar[i], GF_WORKER_WARPS);
    }
    mbar_init(&zbar, GF_WORKER_WARPS);
    mbar_init(&ybar, GF_WORKER_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // TMEM columns: Z (A operand of GEMM-1) [0, 2C), U buffers [2C, 2C + 128), Y [2C + 128, 3C + 128)   (<= 416 of 512)
  const uint32_t tmem_u = tmem + 2 * C;
  const uint32_t tmem_y = tmem_u + 128;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_slices = my_tiles * NS;

  if (warp == GF_WORKER_WARPS) {
    // ================================================================ issuer warp: TMA weight ring + every MMA
    const uint8_t* blob = reinterpret_cast<const uint8_t*>(p.wblob);
    const uint32_t idesc1 = make_idesc_bf16(128, 32), idesc2 = make_idesc_bf16(128, C);
    const uint32_t g_base = smem_u32(smem + L::OFF_G);
    const uint32_t win_base = smem_u32(smem + L::OFF_WIN), wo_base = smem_u32(smem + L::OFF_WO);
    // descriptor templates: everything but the 14-bit start address
    const uint64_t dz_t = make_sdesc(0, 128, L::SBOZ), dg_t = make_sdesc(0, 128, GF_G_SBO), dw_t = make_sdesc(0, 128, 256);
    auto load_slice = [&](int n) {
      const uint8_t* src = blob + (size_t)(n % NS) * L::SLICE;
      const int ws = n % GF_WIN_SLOTS;
      mbar_arrive_expect_tx(&winbar[ws], L::WIN);
      bulk_g2s(smem + L::OFF_WIN + ws * L::WIN, src, L::WIN, &winbar[ws]);
      mbar_arrive_expect_tx(&wobar[n & 3], L::WO);
      bulk_g2s(smem + L::OFF_WO + (n & 3) * L::WO, src + L::WIN, L::WO, &wobar[n & 3]);
    };
    auto issue_gemm1 = [&](int n) {                   // U[n & 1] = Z . W_in[slice n]^T, both 128-row tiles
      const uint32_t wb = (win_base + (uint32_t)(n % GF_WIN_SLOTS) * L::WIN) >> 4;
      if (!(p.debug & 1)) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t d = tmem_u + (uint32_t)(n & 1) * 64 + mt * 32;
          const uint32_t ah = tmem + mt * C, al = ah + C / 2;        // A operand in TMEM: 8 columns per k16 step
#pragma unroll
          for (int ks = 0; ks < C / 16; ++ks) {
            const uint64_t dbh = dz_t | (uint64_t)((wb + ks * 16) & 0x3FFFu);
            const uint64_t dbl = dz_t | (uint64_t)((wb + (4 * L::SBOZ >> 4) + ks * 16) & 0x3FFFu);
            tc_mma_bf16_ts(d, ah + ks * 8, dbh, idesc1, ks == 0 ? 0u : 1u);
            tc_mma_bf16_ts(d, al + ks * 8, dbh, idesc1, 1u);
            tc_mma_bf16_ts(d, ah + ks * 8, dbl, idesc1, 1u);
          }
        }
      }
      tc_commit(&ubar[n & 1]);
    };
    auto issue_gemm2 = [&](int n, bool first) {       // Y (+)= g[n & 1] . W_out[:, slice n]^T   (K = 16)
      const uint32_t gh = (g_base + (uint32_t)(n & 1) * 2 * GF_G_TILE) >> 4, gl = gh + (GF_G_TILE >> 4);
      const uint32_t wh = (wo_base + (uint32_t)(n & 3) * L::WO + GF_DW_BYTES) >> 4, wl = wh + (L::WOUT_T >> 4);
      if (!(p.debug & 1)) {
        const uint64_t dgh = dg_t | (uint64_t)(gh & 0x3FFFu), dgl = dg_t | (uint64_t)(gl & 0x3FFFu);
        const uint64_t dwh = dw_t | (uint64_t)(wh & 0x3FFFu), dwl = dw_t | (uint64_t)(wl & 0x3FFFu);
        tc_mma_bf16(tmem_y, dgh, dwh, idesc2, first ? 0u : 1u);
        tc_mma_bf16(tmem_y, dgl, dwh, idesc2, 1u);
        tc_mma_bf16(tmem_y, dgh, dwl, idesc2, 1u);
      }
      tc_commit(&gbar[n & 1]);
    };
    if (lane == 0) {
      if (total_slices > 0) load_slice(0);
      if (total_slices > 1) load_slice(1);
    }
    int n = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      mbar_wait(&zbar, (uint32_t)ti & 1);                          // Z of this tile is in tensor memory
      for (int s = 0; s < NS; ++s, ++n) {
        // ---- GEMM-1(n): needs W_in(n) and the U buffer n&1 drained (slice n-2)
        if (!(p.debug & 8) || n < 2) mbar_wait(&winbar[n % GF_WIN_SLOTS], (uint32_t)(n / GF_WIN_SLOTS) & 1);
        if (n >= 2) mbar_wait(&dbar[n & 1], (uint32_t)((n - 2) >> 1) & 1);
        tc_fence_after();
        if (elect_one()) issue_gemm1(n);
        __syncwarp();
        // ---- weight prefetch for slice n+2: W_in slot of slice n-1 (GEMM-1(n-1) done), W_out slot of slice n-2
        if (n + 2 < total_slices && !(p.debug & 8)) {
          if (n >= 1) mbar_wait(&ubar[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1);
          if (n >= 2) mbar_wait(&gbar[n & 1], (uint32_t)((n - 2) >> 1) & 1);
          if (lane == 0) load_slice(n + 2);
          __syncwarp();
        }
        // ---- GEMM-2(n-1): needs g(n-1); the first one of a tile overwrites Y, which the epilogue must have drained
        if (s > 0) {
          mbar_wait(&sbar[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1);
          if (s == 1 && ti > 0) mbar_wait(&ybar, (uint32_t)(ti - 1) & 1);
          tc_fence_after();
          if (elect_one()) issue_gemm2(n - 1, s == 1);
          __syncwarp();
        }
      }
      mbar_wait(&sbar[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1);
      if (NS == 1 && ti > 0) mbar_wait(&ybar, (uint32_t)(ti - 1) & 1);
      tc_fence_after();
      if (elect_one()) issue_gemm2(n - 1, NS == 1);
      __syncwarp();
    }
  } else {
    // ================================================================ 16 worker warps
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

    int n = 0, ti = 0;
    if (blockIdx.x < total_tiles) produce_z(blockIdx.x);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const int b = t / tiles_per_img, tr = t - b * tiles_per_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const int y0 = ty * GF_TH, x0 = tx * GF_TW;
      const float* xb = p.x + (size_t)b * p.x_bs;
      for (int s = 0; s < NS; ++s, ++n) {
        mbar_wait(&ubar[n & 1], (uint32_t)(n >> 1) & 1);         // GEMM-1(n) done: U[n & 1] ready
        tc_fence_after();
        // ---- drain U(n): TMEM -> registers, release the TMEM buffer, then registers -> shared [ch][row][col]
        {
          const int q = warp & 3, mt = (warp >> 2) & 1, ch0 = (warp >> 3) * 16;
          const int hp = mt * 128 + q * 32 + lane;
          uint32_t r[16];
          tmem_ld16_nowait(tmem_u + ((uint32_t)(q * 32) << 16) + (uint32_t)(n & 1) * 64 + mt * 32 + ch0, r);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&dbar[n & 1]);
          worker_sync();                                           // stencil(n-1) has finished reading Usm
          if (hp < GF_NHP && !(p.debug & 4)) {
            const int hy = hp / GF_HW, hx = hp - hy * GF_HW;
            float* up = Usm + ch0 * GF_CS + hy * GF_RS + hx;
#pragma unroll
            for (int i = 0; i < 16; ++i) up[i * GF_CS] = __uint_as_float(r[i]);
            if (p.save_u && hy >= 1 && hy <= GF_TH && hx >= 1 && hx <= GF_TW) {
              // channel of row i: a-part (ch0 == 0): s*16 + i ; b-part: hid + s*16 + i
              float* su = p.save_u + (size_t)b * p.u_bs + (size_t)(y0 + hy - 1) * W + (x0 + hx - 1);
              const int cbase = s * GF_HS + (ch0 ? hid : 0);
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (s * GF_HS + i < hid) su[(size_t)(cbase + i) * HWp] = __uint_as_float(r[i]);
            }
          }
        }
        worker_sync();                                             // Usm complete
        if (!(p.debug & 8) || n < 2) mbar_wait(&wobar[n & 3], (uint32_t)(n >> 2) & 1);   // depthwise taps of slice n
        // (g slot n&1 is free: GEMM-2(n-2) was issued before GEMM-1(n), whose completion ubar[n&1] signalled -- MMAs
        //  retire in issue order, so no separate wait on gbar is needed here)
        // ---- stencil + gate.  warp = (core row r, pair half jh); lane = (pair jj of the half, 4-pixel strip xq):
        //      a quarter-warp reads 8 different channels at one strip (conflict-free 16-byte accesses, GF_CS) and the
        //      8 pairs of a lane group are the 8 contiguous k values of one operand row (conflict-free 2-byte stores)
        if (!(p.debug & 2)) {
          const int r = warp >> 1, j = (warp & 1) * 8 + (lane & 7), xq = lane >> 3;
          const float* wdw = reinterpret_cast<const float*>(smem + L::OFF_WO + (n & 3) * L::WO);
          float wv[20];                    // taps of the pair: a0..a8, b0..b8 (+2 pad) as five 16-byte loads
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float4 t4 = lds128(wdw + j * 20 + 4 * i);
            wv[4 * i] = t4.x; wv[4 * i + 1] = t4.y; wv[4 * i + 2] = t4.z; wv[4 * i + 3] = t4.w;
          }
          const float* wa = wv;
          const float* wb = wv + 9;
          float a[4] = {0.f, 0.f, 0.f, 0.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};
          const float* ua = Usm + j * GF_CS + r * GF_RS + 4 * xq;
          const float* ub = ua + 16 * GF_CS;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            // columns 4xq .. 4xq+7 of the halo row (6 are used; the row stride of 20 keeps the second access in the row)
            const float4 a4 = lds128(ua + dy * GF_RS), a2 = lds128(ua + dy * GF_RS + 4);
            const float4 b4 = lds128(ub + dy * GF_RS), b2 = lds128(ub + dy * GF_RS + 4);
            const float va[6] = {a4.x, a4.y, a4.z, a4.w, a2.x, a2.y};
            const float vb[6] = {b4.x, b4.y, b4.z, b4.w, b2.x, b2.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                a[i] = fmaf(wa[dy * 3 + dx], va[i + dx], a[i]);
                bb[i] = fmaf(wb[dy * 3 + dx], vb[i + dx], bb[i]);
              }
          }
          float g[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) g[i] = gelu_fast(a[i]) * bb[i];
          uint8_t* gh = smem + L::OFF_G + (n & 1) * 2 * GF_G_TILE + (j >> 3) * 128 + (j & 7) * 2;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int pr = r * GF_TW + 4 * xq + i;
            const __nv_bfloat16 hi = __float2bfloat16_rn(g[i]);
            const __nv_bfloat16 lo = __float2bfloat16_rn(g[i] - __bfloat162float(hi));
            const uint32_t off = (uint32_t)(pr >> 3) * GF_G_SBO + (uint32_t)(pr & 7) * 16;
            *reinterpret_cast<__nv_bfloat16*>(gh + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(gh + GF_G_TILE + off) = lo;
          }
          if (p.save_g && s * GF_HS + j < hid) {
            float* sg = p.save_g + (size_t)b * p.g_bs + (size_t)(s * GF_HS + j) * HWp + (size_t)(y0 + r) * W + x0 + 4 * xq;
            *reinterpret_cast<float4*>(sg) = make_float4(g[0], g[1], g[2], g[3]);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sbar[n & 1]);
      }
      // ---- every GEMM-1 of this tile has completed (its U was drained): Z may be rebuilt for the next tile while
      //      the tensor pipe finishes GEMM-2 of the last slice
      if (t + (int)gridDim.x < total_tiles) produce_z(t + gridDim.x);
      // ---- epilogue: Y (+ x) -> HBM, LayerNorm statistics of y
      mbar_wait(&gbar[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1);
      tc_fence_after();
      {
        const int q = warp & 3, cg = warp >> 2;
        const int pr = q * 32 + lane, r = pr >> 4, cx = pr & 15;
        const size_t pix = (size_t)(y0 + r) * W + x0 + cx;
        const float* xr = xb + pix;
        float* yo = p.y + (size_t)b * p.y_bs + pix;
        const bool want_stats = p.stats_out != nullptr;
        const float shift = want_stats ? __ldg(xr) : 0.f;
        float s1 = 0.f, s2 = 0.f;
        for (int gi = cg; gi < C / 8; gi += 4) {
          uint32_t rr[8];
          tmem_ld8_nowait(tmem_y + ((uint32_t)(q * 32) << 16) + gi * 8, rr);
          float res[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) res[i] = p.residual ? __ldg(xr + (size_t)(gi * 8 + i) * HWp) : 0.f;
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float yv = __uint_as_float(rr[i]) + res[i];
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
          // (stp is rewritten by the next tile's epilogue only after 2 * NS worker barriers)
        }
      }
    }
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

template <int C>
static int launch_gdfn(const rcot_gdfn_params& p, cudaStream_t stream) {
  using L = GfLayout<C>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gdfn_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
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
  gdfn_fwd_kernel<C><<<grid, GF_THREADS, L::TOTAL, stream>>>(p, tiles_x, tpi, (int)total);
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
  RCOT_REQUIRE(p.B > 0 && p.hid > 0, "gdfn_fwd: bad sizes");
  RCOT_REQUIRE(rcot_gdfn_supported(p.C, p.H, p.W), "gdfn_fwd: needs C in {48, 96}, H %% 8 == 0, W %% 16 == 0 (got C=%d %dx%d)",
               p.C, p.H, p.W);
  if (p.ln_stats) RCOT_REQUIRE(p.ln_gamma && p.ln_beta, "gdfn_fwd: LayerNorm needs gamma and beta");
  RCOT_REQUIRE((long)p.B * (p.H / GF_TH) * (p.W / GF_TW) < (1L << 31), "gdfn_fwd: too many tiles");
  if (p.save_g) RCOT_REQUIRE((reinterpret_cast<uintptr_t>(p.save_g) & 15) == 0 && p.g_bs % 4 == 0, "gdfn_fwd: save_g alignment");
  return p.C == 48 ? launch_gdfn<48>(p, (cudaStream_t)st) : launch_gdfn<96>(p, (cudaStream_t)st);
}
