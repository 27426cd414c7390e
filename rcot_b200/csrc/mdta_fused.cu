// mdta_fused.cu -- MDTA phase 1 as ONE kernel (Net_Restormer.py:29-41 with the block's LayerNorm):
//     pre = W_qkv . LN(x)            1x1 conv, 3C channels          (never reaches HBM)
//     [q; k; v] = dw3x3(pre)         depthwise, zero padded         (q, k never reach HBM)
//     G   += q k^T  per image (all heads at once: the diagonal c x c blocks are the heads' Grams)
//     ssq += row sums of q^2 and k^2 (the L2 norms of F.normalize)
//     v   -> HBM                     (the only large output: phase 2 is y = x + M v, one pm_gemm launch)
// Same tile machinery as the one-kernel GDFN forward (gdfn_fused.cu, version 6): 8 x 16-pixel tiles + 1-pixel halo,
// LN(x) of the halo tile resident in TENSOR memory as the A operand, the 3C channels walked in slices of 32,
//   ISSUER 1 : W_qkv slice ring (TMA) + GEMM-1  U_n[256 x 32] = Z . W_qkv[slice]^T  (A from TMEM)
//   4 DRAIN warps: U_n TMEM -> shared memory [channel][10][18] (3 buffers)
//   16 STENCIL warps in two groups on alternate slices, 2 x 4 pixels x 2 channels per thread:
//        v slices  -> 16-byte global stores
//        q/k slices-> bf16 hi/lo operand rows [channel x 128 pixels] in shared memory (K-major, K = the tile's pixels)
//                     + per-channel sums of squares (warp shuffles, shared-memory atomics)
//   ISSUER 2 : once per tile, G[128 x C] += Q . K^T on tcgen05 (24 MMAs), accumulated in TMEM ACROSS the tiles of an
//              image: a CTA owns a CONTIGUOUS run of tiles, so the Gram leaves the SM only when the image changes
//              (red.global.add into the zero-initialised G: at most two CTAs share an image boundary).
// The v slices of a tile are processed first, so the Gram MMAs of the previous tile retire long before the q/k operand
// rows are overwritten.  Reduction order of G: pixels of a tile inside one MMA chain (fp32 accumulate in TMEM, bf16x3
// split products hi*hi + lo*hi + hi*lo), tiles in raster order per CTA, CTAs by atomics -- fp32-class, order fixed up to
// the (at most one) atomic merge per image.
// Algorithmic HBM bytes: read x (+ halo re-reads, L2 hits) + write v = 2 C per pixel instead of ~12 C for the three
// launches (pm_gemm x->pre, dw_plain, pk_gemm Gram).  Optional outputs pre / q,k feed the existing backward.
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "fused_tile.cuh"
#include "tc.cuh"

namespace rcot {

constexpr int MF_S_WARPS = 16;                       // stencil / Z / flush warps: two groups of 8
constexpr int MF_D_WARPS = 4;                        // drain warps (TMEM lane quarter = warp & 3)
constexpr int MF_ISSUER = MF_S_WARPS + MF_D_WARPS;   // GEMM-1 issuer; MF_ISSUER + 1 issues the Gram MMAs
constexpr int MF_WARPS = MF_ISSUER + 2;
constexpr int MF_THREADS = MF_WARPS * 32;            // 704
constexpr int MF_U_BUFS = 3;
constexpr int MF_W_SLOTS = 3;
constexpr uint32_t MF_U_BYTES = 32 * GF_CS * 4;
constexpr uint32_t MF_TAP_BYTES = 16 * 20 * sizeof(float);   // per channel pair (j, j + 16): 9 + 9 taps, 2 pad
constexpr uint32_t MF_QK_SBO = 16 * 128;             // q / k operand [C x 128 pixels]: 8-row group stride, LBO 128

template <int C>
struct MfLayout {
  static constexpr int NCH = 3 * C;
  static constexpr int NS = (NCH + 31) / 32;                   // channel slices (C = 48: the last one is half empty)
  static constexpr int SV = 2 * C / 32;                        // first v slice; processing order: v slices, then q, k
  static constexpr uint32_t SBOZ = (C / 8) * 128;              // W_qkv operand: [32 x C], LBO 128
  static constexpr uint32_t WIN = 2 * 4 * SBOZ;                // 2 terms
  static constexpr uint32_t QK_T = (C / 8) * MF_QK_SBO;        // one term of Q or K
  static constexpr uint32_t OFF_U = 0;
  static constexpr uint32_t OFF_WIN = OFF_U + MF_U_BUFS * MF_U_BYTES;
  static constexpr uint32_t OFF_QK = OFF_WIN + MF_W_SLOTS * WIN;   // Q hi, Q lo, K hi, K lo
  static constexpr uint32_t OFF_GB = OFF_QK + 4 * QK_T;        // gamma, beta
  static constexpr uint32_t OFF_SS = OFF_GB + 2 * C * 4;       // sums of squares of the current image [2C]
  static constexpr uint32_t TOTAL = OFF_SS + 2 * C * 4;
  static constexpr size_t BLOB = (size_t)NS * (WIN + MF_TAP_BYTES);
  static_assert((2 * C) % 32 == 0, "v must start on a slice boundary");
  static_assert(QK_T + 16 * MF_QK_SBO <= 4 * QK_T, "the M = 128 A operand read of Q lo must stay inside the q/k buffers");
};

__device__ __forceinline__ float4 ldg128(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

template <int C>
__global__ void __launch_bounds__(MF_THREADS, 1) mdta_p1_kernel(const rcot_mdta_p1_params p, const int tiles_x,
                                                               const int tiles_per_img, const int total_tiles) {
  using L = MfLayout<C>;
  constexpr int NS = L::NS;
  extern __shared__ __align__(128) uint8_t smem[];
  // TMA: winbar.  tcgen05.commit: ubar (GEMM-1 retired), grambar (Gram MMAs of a tile retired).
  // drain -> issuer 1: dbar.  drain -> stencil: ufull.  stencil -> drain: uempty.
  // stencil -> issuer 1: zbar (Z in TMEM).  stencil -> issuer 2: qkbar (q/k rows of the tile written), flushbar.
  __shared__ uint64_t winbar[MF_W_SLOTS], ubar[2], dbar[2], ufull[MF_U_BUFS], uempty[MF_U_BUFS], zbar, qkbar, grambar, flushbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, W = p.W, HWp = H * W;
  const bool LN = p.ln_stats != nullptr;
  float* gb = reinterpret_cast<float*>(smem + L::OFF_GB);
  float* ss = reinterpret_cast<float*>(smem + L::OFF_SS);

  if (LN)
    for (int c = tid; c < C; c += MF_THREADS) {
      gb[c] = __ldg(p.ln_gamma + c);
      gb[C + c] = __ldg(p.ln_beta + c);
    }
  for (int c = tid; c < 2 * C; c += MF_THREADS) ss[c] = 0.f;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < MF_W_SLOTS; ++i) mbar_init(&winbar[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ubar[i], 1);
      mbar_init(&dbar[i], MF_D_WARPS);
    }
    for (int i = 0; i < MF_U_BUFS; ++i) {
      mbar_init(&ufull[i], MF_D_WARPS);
      mbar_init(&uempty[i], MF_S_WARPS / 2);
    }
    mbar_init(&zbar, MF_S_WARPS);
    mbar_init(&qkbar, MF_S_WARPS);
    mbar_init(&grambar, 1);
    mbar_init(&flushbar, MF_S_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // TMEM columns: Z [0, 2C), U buffers [2C, 2C + 128), Gram [2C + 128, 3C + 128)
  const uint32_t tmem_u = tmem + 2 * C;
  const uint32_t tmem_g = tmem_u + 128;
  // contiguous run of tiles (raster order inside an image, images in order)
  const int t0 = (int)((long long)total_tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)total_tiles * (blockIdx.x + 1) / gridDim.x);
  const int my_tiles = t1 - t0;
  const int total_slices = my_tiles * NS;
  auto flushes_after = [&](int t) { return t + 1 == t1 || (t + 1) / tiles_per_img != t / tiles_per_img; };

  if (warp == MF_ISSUER) {
    // ================================================================ issuer 1: W_qkv ring + GEMM-1
    const uint8_t* blob = reinterpret_cast<const uint8_t*>(p.wblob);
    const uint32_t idesc1 = make_idesc_bf16(128, 32);
    const uint32_t win_base = smem_u32(smem + L::OFF_WIN);
    const uint64_t dz_t = make_sdesc(0, 128, L::SBOZ);
    auto load_win = [&](int n, int s) {                           // s = processing index inside the tile
      const int ws = n % MF_W_SLOTS;
      const int sl = (s + L::SV) % NS;
      mbar_arrive_expect_tx(&winbar[ws], L::WIN);
      bulk_g2s(smem + L::OFF_WIN + ws * L::WIN, blob + (size_t)sl * L::WIN, L::WIN, &winbar[ws]);
    };
    if (lane == 0) {
      if (total_slices > 0) load_win(0, 0);
      if (total_slices > 1) load_win(1, 1 % NS);
    }
    __syncwarp();
    int n = 0, s2 = 2 % NS;                                         // s2 = processing index of global slice n + 2
    for (int ti = 0; ti < my_tiles; ++ti) {
      mbar_wait(&zbar, (uint32_t)ti & 1);
      for (int s = 0; s < NS; ++s, ++n) {
        mbar_wait(&winbar[n % MF_W_SLOTS], (uint32_t)(n / MF_W_SLOTS) & 1);
        if (n >= 2) mbar_wait(&dbar[n & 1], (uint32_t)((n - 2) >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t wb = (win_base + (uint32_t)(n % MF_W_SLOTS) * L::WIN) >> 4;
          if (!(p.debug & 1)) {
            const uint32_t d0 = tmem_u + (uint32_t)(n & 1) * 64, d1 = d0 + 32;
            const uint32_t ah0 = tmem, al0 = ah0 + C / 2, ah1 = tmem + C, al1 = ah1 + C / 2;
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
        // W_qkv of slice n+2 into the slot of slice n-1, once GEMM-1(n-1) has retired (it is ahead of GEMM-1(n) in
        // the pipe, so the tensor cores stay busy while this warp waits)
        if (n + 2 < total_slices) {
          if (n >= 1) mbar_wait(&ubar[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1);
          if (lane == 0) load_win(n + 2, s2);
          __syncwarp();
        }
        s2 = (s2 + 1 == NS) ? 0 : s2 + 1;
      }
    }
  } else if (warp == MF_ISSUER + 1) {
    // ================================================================ issuer 2: G += Q . K^T, once per tile
    const uint32_t idesc = make_idesc_bf16(128, C);
    const uint32_t qh = smem_u32(smem + L::OFF_QK) >> 4, ql = qh + (L::QK_T >> 4);
    const uint32_t kh = qh + (2 * L::QK_T >> 4), kl = kh + (L::QK_T >> 4);
    const uint64_t d_t = make_sdesc(0, 128, MF_QK_SBO);
    int nflush = 0;
    bool fresh = true;                                               // the accumulator holds nothing of this image yet
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int t = t0 + ti;
      mbar_wait(&qkbar, (uint32_t)ti & 1);                           // every q / k row of the tile is in shared memory
      if (fresh && nflush > 0) mbar_wait(&flushbar, (uint32_t)(nflush - 1) & 1);   // the old image's Gram was read out
      tc_fence_after();
      if (elect_one()) {
        if (!(p.debug & 1)) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                           // 128 pixels = 8 k16 steps of two core matrices
            const uint64_t aqh = d_t | (uint64_t)((qh + ks * 16) & 0x3FFFu), aql = d_t | (uint64_t)((ql + ks * 16) & 0x3FFFu);
            const uint64_t bkh = d_t | (uint64_t)((kh + ks * 16) & 0x3FFFu), bkl = d_t | (uint64_t)((kl + ks * 16) & 0x3FFFu);
            tc_mma_bf16(tmem_g, aqh, bkh, idesc, (fresh && ks == 0) ? 0u : 1u);
            tc_mma_bf16(tmem_g, aql, bkh, idesc, 1u);
            tc_mma_bf16(tmem_g, aqh, bkl, idesc, 1u);
          }
        }
        tc_commit(&grambar);
      }
      __syncwarp();
      fresh = flushes_after(t);
      if (fresh) ++nflush;
    }
  } else if (warp >= MF_S_WARPS) {
    // ================================================================ 4 drain warps: U_n  TMEM -> registers -> shared
    const int q = warp & 3;
    const int hp0 = q * 32 + lane, hp1 = 128 + hp0;
    const int hy0 = hp0 / GF_HW, hx0 = hp0 - hy0 * GF_HW;
    const int hy1 = hp1 / GF_HW, hx1 = hp1 - hy1 * GF_HW;
    const bool two = q < 2;
    int n = 0;
    for (int t = t0; t < t1; ++t) {
      if (t + 1 < t1) {
        // pull the x rows of the NEXT tile's halo region towards L2
        const int t2 = t + 1;
        const int b2 = t2 / tiles_per_img, tr2 = t2 - b2 * tiles_per_img;
        const int ty2 = tr2 / tiles_x, tx2 = tr2 - ty2 * tiles_x;
        const int gx0 = max(tx2 * GF_TW - 1, 0);
        const float* xn = p.x + (size_t)b2 * p.x_bs + gx0;
        for (int idx = (warp - MF_S_WARPS) * 32 + lane; idx < C * GF_HH; idx += MF_D_WARPS * 32) {
          const int ch = idx / GF_HH, gy = ty2 * GF_TH - 1 + (idx - ch * GF_HH);
          if ((unsigned)gy < (unsigned)H) {
            const float* a0 = xn + (size_t)ch * HWp + (size_t)gy * W;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + 16));
          }
        }
      }
      for (int s = 0; s < NS; ++s, ++n) {
        mbar_wait(&ubar[n & 1], (uint32_t)(n >> 1) & 1);
        tc_fence_after();
        const uint32_t ta = tmem_u + ((uint32_t)(q * 32) << 16) + (uint32_t)(n & 1) * 64;
        const int ub = n % MF_U_BUFS;
        float* Ub = reinterpret_cast<float*>(smem + L::OFF_U + ub * MF_U_BYTES);
        auto put = [&](const uint32_t (&r)[32], int hp, int hy, int hx) {
          if (hp >= GF_NHP || (p.debug & 4)) return;
          float* up = Ub + hy * GF_RS + hx;
#pragma unroll
          for (int i = 0; i < 32; ++i) up[i * GF_CS] = __uint_as_float(r[i]);
        };
        uint32_t r[32];
        tmem_ld32_nowait(ta, r);
        tmem_ld_wait();
        if (n >= MF_U_BUFS) mbar_wait(&uempty[ub], (uint32_t)(n / MF_U_BUFS - 1) & 1);
        put(r, hp0, hy0, hx0);
        if (two) {
          tmem_ld32_nowait(ta + 32, r);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dbar[n & 1]);
        if (two) put(r, hp1, hy1, hx1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&ufull[ub]);
      }
    }
  } else {
    // ================================================================ 16 stencil warps (two groups of 8)
    auto produce_z = [&](int t) {
      const int b = t / tiles_per_img, tr = t - b * tiles_per_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const int y0 = ty * GF_TH, x0 = tx * GF_TW;
      const int hp = tid & 255, half = tid >> 8;
      const int hy = hp / GF_HW, hx = hp - hy * GF_HW;
      const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
      const bool inimg = hp < GF_NHP && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
      const int mt = hp >> 7;
      const uint32_t zaddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * C);
      constexpr int NG = C / 16;
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
      } else {                                         // outside the image: zero rows = the conv's zero padding of pre
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) v[g][i] = 0.f;
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
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
    const int rp = wg >> 1, j = (wg & 1) * 8 + (lane & 7), xq = lane >> 3;
    const float* taps = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.wblob) + (size_t)NS * L::WIN);
    int n = 0, ti = 0;
    if (my_tiles > 0) produce_z(t0);
    for (int t = t0; t < t1; ++t, ++ti) {
      const int b = t / tiles_per_img, tr = t - b * tiles_per_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const int y0 = ty * GF_TH, x0 = tx * GF_TW;
      const int nl = n + NS - 1;
      bool qk_free = (ti == 0);                                      // q / k operand rows may be overwritten
      for (int s = 0; s < NS; ++s, ++n) {
        if ((n & 1) != grp) continue;
        if (s >= NS - 2 && t + 1 < t1) {
          // this group's last slice of the tile: once GEMM-1(nl) has retired nothing reads Z any more -- build the next
          // tile's Z now (see gdfn_fused.cu for why the ufull(nl) peek cannot alias)
          mbar_wait(&ufull[nl % MF_U_BUFS], (uint32_t)(nl / MF_U_BUFS) & 1);
          tc_fence_after();
          produce_z(t + 1);
        }
        const int ub = n % MF_U_BUFS;
        const int sl = (s + L::SV) % NS;                             // channel slice: channels 32 sl .. 32 sl + 31
        const int ca = 32 * sl + j, cb = ca + 16;                    // this thread's two channels
        const bool is_v = sl >= L::SV;
        float wv[20];                                                // taps: a0..a8, b0..b8 (+2 pad)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const float4 t4 = ldg128(taps + (size_t)(sl * 16 + j) * 20 + 4 * i);
          wv[4 * i] = t4.x; wv[4 * i + 1] = t4.y; wv[4 * i + 2] = t4.z; wv[4 * i + 3] = t4.w;
        }
        mbar_wait(&ufull[ub], (uint32_t)(n / MF_U_BUFS) & 1);
        if (!is_v && !qk_free) {
          mbar_wait(&grambar, (uint32_t)(ti - 1) & 1);               // the previous tile's Gram MMAs have retired
          qk_free = true;
        }
        if (!(p.debug & 2)) {
          const float* wa = wv;
          const float* wb = wv + 9;
          float a[2][4], bb[2][4];
#pragma unroll
          for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int i = 0; i < 4; ++i) a[o][i] = bb[o][i] = 0.f;
          const float* ua = reinterpret_cast<const float*>(smem + L::OFF_U + ub * MF_U_BYTES) + j * GF_CS +
                            (2 * rp) * GF_RS + 4 * xq;
          const float* ubp = ua + 16 * GF_CS;
#pragma unroll
          for (int hr = 0; hr < 4; ++hr) {
            const float4 a4 = lds128(ua + hr * GF_RS), a2 = lds128(ua + hr * GF_RS + 4);
            const float4 b4 = lds128(ubp + hr * GF_RS), b2 = lds128(ubp + hr * GF_RS + 4);
            const float va[6] = {a4.x, a4.y, a4.z, a4.w, a2.x, a2.y};
            const float vb[6] = {b4.x, b4.y, b4.z, b4.w, b2.x, b2.y};
            if (p.save_pre && (hr == 1 || hr == 2)) {
              float* sp = p.save_pre + (size_t)b * p.pre_bs + (size_t)(y0 + 2 * rp + hr - 1) * W + x0 + 4 * xq;
              if (ca < L::NCH) *reinterpret_cast<float4*>(sp + (size_t)ca * HWp) = make_float4(va[1], va[2], va[3], va[4]);
              if (cb < L::NCH) *reinterpret_cast<float4*>(sp + (size_t)cb * HWp) = make_float4(vb[1], vb[2], vb[3], vb[4]);
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
          if (is_v) {
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              float* vp = p.v + (size_t)b * p.v_bs + (size_t)(y0 + 2 * rp + o) * W + x0 + 4 * xq;
              if (ca < L::NCH) *reinterpret_cast<float4*>(vp + (size_t)(ca - 2 * C) * HWp) = make_float4(a[o][0], a[o][1], a[o][2], a[o][3]);
              if (cb < L::NCH) *reinterpret_cast<float4*>(vp + (size_t)(cb - 2 * C) * HWp) = make_float4(bb[o][0], bb[o][1], bb[o][2], bb[o][3]);
            }
          } else {
            // q / k: operand rows (row = channel inside q resp. k, K = pixel of the tile) + sums of squares
            float sa = 0.f, sb = 0.f;
            uint8_t* oa = smem + L::OFF_QK + (ca < C ? 0u : 2 * L::QK_T) + (uint32_t)((ca < C ? ca : ca - C) >> 3) * MF_QK_SBO +
                          (uint32_t)(ca & 7) * 16 + (uint32_t)(xq & 1) * 8;
            uint8_t* ob = smem + L::OFF_QK + (cb < C ? 0u : 2 * L::QK_T) + (uint32_t)((cb < C ? cb : cb - C) >> 3) * MF_QK_SBO +
                          (uint32_t)(cb & 7) * 16 + (uint32_t)(xq & 1) * 8;
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              const uint32_t koff = (uint32_t)((2 * rp + o) * 2 + (xq >> 1)) * 128;   // (k / 8) * LBO, k = (2rp+o)*16 + 4xq + i
              uint32_t h2[2], l2[2];
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const __nv_bfloat162 hp2 = __floats2bfloat162_rn(a[o][2 * i], a[o][2 * i + 1]);
                const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hp2);
                h2[i] = hb;
                const __nv_bfloat162 lp2 = __floats2bfloat162_rn(a[o][2 * i] - __uint_as_float(hb << 16),
                                                                 a[o][2 * i + 1] - __uint_as_float(hb & 0xffff0000u));
                l2[i] = *reinterpret_cast<const uint32_t*>(&lp2);
                sa = fmaf(a[o][2 * i], a[o][2 * i], sa);
                sa = fmaf(a[o][2 * i + 1], a[o][2 * i + 1], sa);
              }
              *reinterpret_cast<uint2*>(oa + koff) = make_uint2(h2[0], h2[1]);
              *reinterpret_cast<uint2*>(oa + L::QK_T + koff) = make_uint2(l2[0], l2[1]);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const __nv_bfloat162 hp2 = __floats2bfloat162_rn(bb[o][2 * i], bb[o][2 * i + 1]);
                const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hp2);
                h2[i] = hb;
                const __nv_bfloat162 lp2 = __floats2bfloat162_rn(bb[o][2 * i] - __uint_as_float(hb << 16),
                                                                 bb[o][2 * i + 1] - __uint_as_float(hb & 0xffff0000u));
                l2[i] = *reinterpret_cast<const uint32_t*>(&lp2);
                sb = fmaf(bb[o][2 * i], bb[o][2 * i], sb);
                sb = fmaf(bb[o][2 * i + 1], bb[o][2 * i + 1], sb);
              }
              *reinterpret_cast<uint2*>(ob + koff) = make_uint2(h2[0], h2[1]);
              *reinterpret_cast<uint2*>(ob + L::QK_T + koff) = make_uint2(l2[0], l2[1]);
              if (p.save_qk) {
                float* qp = p.save_qk + (size_t)b * p.qk_bs + (size_t)(y0 + 2 * rp + o) * W + x0 + 4 * xq;
                *reinterpret_cast<float4*>(qp + (size_t)ca * HWp) = make_float4(a[o][0], a[o][1], a[o][2], a[o][3]);
                *reinterpret_cast<float4*>(qp + (size_t)cb * HWp) = make_float4(bb[o][0], bb[o][1], bb[o][2], bb[o][3]);
              }
            }
            // the four strips of a channel sit in lanes j, j+8, j+16, j+24
            sa += __shfl_xor_sync(0xffffffffu, sa, 8);
            sb += __shfl_xor_sync(0xffffffffu, sb, 8);
            sa += __shfl_xor_sync(0xffffffffu, sa, 16);
            sb += __shfl_xor_sync(0xffffffffu, sb, 16);
            if (xq == 0) {
              atomicAdd(&ss[ca], sa);
              atomicAdd(&ss[cb], sb);
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&uempty[ub]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&qkbar);                           // this warp's q / k rows of the tile are written
      if (flushes_after(t)) {
        // ---- last tile of this image on this CTA: Gram and sums of squares leave the SM
        mbar_wait(&grambar, (uint32_t)ti & 1);
        tc_fence_after();
        worker_sync();                                               // every warp's shared-memory atomics have landed
        const int c = C / p.heads;
        const int qd = warp & 3, cg = warp >> 2;
        const int row = qd * 32 + lane;
        for (int g8 = cg; g8 < C / 8; g8 += 4) {                     // warp-uniform
          uint32_t rr[8];
          tmem_ld8_nowait(tmem_g + ((uint32_t)(qd * 32) << 16) + g8 * 8, rr);
          tmem_ld_wait();
          if (row < C) {
            const int hrow = row / c;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int col = g8 * 8 + i;
              if (col / c == hrow)
                atomicAdd(p.G + (((size_t)b * p.heads + hrow) * c + (row - hrow * c)) * c + (col - hrow * c), __uint_as_float(rr[i]));
            }
          }
        }
        if (tid < 2 * C) {
          atomicAdd(p.sumsq + (size_t)b * 2 * C + tid, ss[tid]);
          ss[tid] = 0.f;
        }
        tc_fence_before();
        worker_sync();
        if (lane == 0) mbar_arrive(&flushbar);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------- weight blob
template <int C>
__global__ void mdta_pack_kernel(const float* __restrict__ w_qkv, const float* __restrict__ w_dw, uint8_t* __restrict__ blob) {
  using L = MfLayout<C>;
  const int sl = blockIdx.x;
  uint8_t* dst = blob + (size_t)sl * L::WIN;
  for (int e = threadIdx.x; e < 32 * C; e += blockDim.x) {
    const int i = e / C, k = e - i * C;
    const int ch = 32 * sl + i;
    const float w = ch < L::NCH ? w_qkv[(size_t)ch * C + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const uint32_t off = (uint32_t)(i >> 3) * L::SBOZ + (uint32_t)(k >> 3) * 128 + (uint32_t)(i & 7) * 16 + (uint32_t)(k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(dst + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(dst + 4 * L::SBOZ + off) = lo;
  }
  float* dw = reinterpret_cast<float*>(blob + (size_t)L::NS * L::WIN + (size_t)sl * MF_TAP_BYTES);
  for (int e = threadIdx.x; e < 16 * 20; e += blockDim.x) {
    const int jp = e / 20, tp = e - jp * 20;           // pair (channel jp, channel jp + 16): slots 0..8, 9..17, pad
    const int ch = 32 * sl + (tp < 9 ? jp : jp + 16);
    float w = 0.f;
    if (tp < 18 && ch < L::NCH) w = w_dw[(size_t)ch * 9 + (tp < 9 ? tp : tp - 9)];
    dw[e] = w;
  }
}

static int mf_num_sms() {
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
static int launch_mdta_p1(const rcot_mdta_p1_params& p, cudaStream_t stream) {
  using L = MfLayout<C>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mdta_p1_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
    if (e != cudaSuccess) {
      set_error("mdta_p1: cudaFuncSetAttribute(%u bytes): %s", L::TOTAL, cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  const int tiles_x = p.W / GF_TW, tiles_y = p.H / GF_TH;
  const int tpi = tiles_x * tiles_y;
  const long total = (long)tpi * p.B;
  const int grid = (int)(total < mf_num_sms() ? total : mf_num_sms());
  mdta_p1_kernel<C><<<grid, MF_THREADS, L::TOTAL, stream>>>(p, tiles_x, tpi, (int)total);
  return check_launch("mdta_p1");
}

}  // namespace rcot

using namespace rcot;

extern "C" int rcot_mdta_p1_supported(int C, int H, int W, int heads) {
  return (C == 48 || C == 96) && H % GF_TH == 0 && W % GF_TW == 0 && H > 0 && W > 0 && heads > 0 && C % heads == 0;
}

extern "C" size_t rcot_mdta_p1_blob_bytes(int C) {
  if (C == 48) return MfLayout<48>::BLOB;
  if (C == 96) return MfLayout<96>::BLOB;
  return 0;
}

extern "C" int rcot_mdta_p1_pack(const float* w_qkv, const float* w_dw, void* blob, int C, rcot_stream_t st) {
  RCOT_REQUIRE(w_qkv && w_dw && blob, "mdta_p1_pack: bad arguments");
  RCOT_REQUIRE(C == 48 || C == 96, "mdta_p1_pack: the fused MDTA kernel is built for C = 48 and 96 (got %d)", C);
  if (C == 48)
    mdta_pack_kernel<48><<<MfLayout<48>::NS, 256, 0, (cudaStream_t)st>>>(w_qkv, w_dw, reinterpret_cast<uint8_t*>(blob));
  else
    mdta_pack_kernel<96><<<MfLayout<96>::NS, 256, 0, (cudaStream_t)st>>>(w_qkv, w_dw, reinterpret_cast<uint8_t*>(blob));
  return check_launch("mdta_p1_pack");
}

extern "C" int rcot_mdta_p1(const rcot_mdta_p1_params* pp, rcot_stream_t st) {
  RCOT_REQUIRE(pp != nullptr, "mdta_p1: null params");
  const rcot_mdta_p1_params& p = *pp;
  RCOT_REQUIRE(p.x && p.v && p.G && p.sumsq && p.wblob, "mdta_p1: null tensor pointer");
  RCOT_REQUIRE(p.B > 0, "mdta_p1: bad sizes");
  RCOT_REQUIRE(rcot_mdta_p1_supported(p.C, p.H, p.W, p.heads),
               "mdta_p1: needs C in {48, 96}, H %% 8 == 0, W %% 16 == 0, C %% heads == 0 (got C=%d %dx%d heads=%d)", p.C, p.H, p.W,
               p.heads);
  if (p.ln_stats) RCOT_REQUIRE(p.ln_gamma && p.ln_beta, "mdta_p1: LayerNorm needs gamma and beta");
  RCOT_REQUIRE((long)p.B * (p.H / GF_TH) * (p.W / GF_TW) < (1L << 31), "mdta_p1: too many tiles");
  RCOT_REQUIRE((reinterpret_cast<uintptr_t>(p.v) & 15) == 0 && p.v_bs % 4 == 0, "mdta_p1: v alignment");
  if (p.save_pre) RCOT_REQUIRE((reinterpret_cast<uintptr_t>(p.save_pre) & 15) == 0 && p.pre_bs % 4 == 0, "mdta_p1: save_pre alignment");
  if (p.save_qk) RCOT_REQUIRE((reinterpret_cast<uintptr_t>(p.save_qk) & 15) == 0 && p.qk_bs % 4 == 0, "mdta_p1: save_qk alignment");
  return p.C == 48 ? launch_mdta_p1<48>(p, (cudaStream_t)st) : launch_mdta_p1<96>(p, (cudaStream_t)st);
}
