// gemm_pm.cu -- "pixel-as-M" implicit GEMM on tcgen05:
//     out[b, n, p] = epilogue( sum_k A(b, p, k) * W[n, k] )
// Persistent, warp-specialised CTAs (one per SM), 128 pixels x BN<=256 output channels per tile:
//   * 8 producer warps gather the A operand straight from the NCHW fp32 activations (two threads per
//     pixel row; optional concat of two tensors, LayerNorm prologue, forward or transposed conv
//     geometry), split fp32 -> bf16 hi/lo and store it in the no-swizzle K-major core-matrix layout;
//     producer thread 0 also brings the pre-packed weight stage in with one cp.async.bulk (TMA engine).
//   * 1 MMA warp: a single thread issues tcgen05.mma (hi*hi + lo*hi + hi*lo for fp32-class accuracy)
//     into one of TWO TMEM accumulator buffers, commits stages back to the producers.
//   * 4 epilogue warps read the finished accumulator with tcgen05.ld (thread = pixel, registers =
//     channels) and write coalesced NCHW rows with bias / LeakyReLU / sign-mask / residual /
//     accumulate fused -- while the MMA warp already works on the next tile in the other buffer.
// Rows are the pixels of the whole batch flattened (b, y, x) unless the weights are per image
// (MDTA's folded matrices), so small images still fill 128-row tiles.
// Replaces the ATen conv2d calls listed in include/rcot_b200.h (rcot_pm_params).
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "tc.cuh"
#include "tmap.cuh"
#include <stdlib.h>

namespace rcot {

constexpr int PM_MAX_STAGES = 6;
constexpr int PM_PROD_WARPS = 8;
constexpr int PM_PROD_THREADS = PM_PROD_WARPS * 32;
constexpr int PM_THREADS = (PM_PROD_WARPS + 1 + 4) * 32;  // producers, MMA warp, 4 epilogue warps
// TMA-staged variant (1x1 convs on feature maps with H*W % 128 == 0): one more warp issues, per K chunk, ONE tensor
// copy (cp.async.bulk.tensor.3d: box = 128 pixels x 32 channels x 1 image of the fp32 NCHW activation, described by
// a CUtensorMap built per launch) into a ring of PM_RAW shared-memory slots; the producers read their 16 values
// from there.  The ring holds 64 KB in flight per SM -- with register prefetch the producers keep 32 KB in flight,
// which caps a 1 us-latency stream at ~60 % of the HBM rate.  (32 separate 512-byte cp.async.bulk copies per chunk
// were measured SLOWER than the register path: the copy engine is request-bound at that size.)
constexpr int PM_RAW = 8;                                       // most slots of the raw ring (g.nraw are used: 4 by default)
constexpr uint32_t PM_RAW_BYTES = KC * 128 * sizeof(float);   // 16 KB: 32 channels x 128 pixels

struct PmGeom {
  int Ktot, nk, BN, passes, stages, tiles_m, tiles_per_img, flat, c1_aligned, nraw;
  // parity classes of a stride-2 transposed conv (dgrad of the 4x4 stride-2 convs): an output pixel only sees the
  // taps with ky = (y + pad) mod 2 (+2), kx likewise -- 4 of 16 -- so tiles hold pixels of ONE (y&1, x&1) class and
  // the K loop runs over that class's 4 taps only (nk = 4 * C1/32 chunks) instead of multiplying 75 % zeros.
  int cls, tpc, nk_full; // cls: enabled; tpc: tiles per class; nk_full: K chunks of the packed weights (nk = the loop's)
  // N slices: when tiles_m * passes cannot fill the SMs (F_net's deep convs: a few hundred pixels, K up to 8192) a
  // pass of BNf packed columns is split into nslice work items of BN = BNf / nslice columns each (their own MMA N,
  // TMEM buffers and weight rows), so 2-4x more CTAs stream the long K loop.  `passes` counts items along N.
  int BNf, nslice;
  long rows_total;
  uint32_t tmem_cols;
};

// ABF / OBF (1x1 TMA variant only): the gather source / the output is a bf16 tensor (bf16-storage mode).
// LNB: LayerNorm-BACKWARD epilogue (1x1, N = C <= 256 = one pass): the accumulator row of a pixel is dz = dL/dLN(x);
// the epilogue turns it into dx = [dy +] rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dz * gamma, and accumulates
// dgamma += sum_p dz * xhat, dbeta += sum_p dz -- the separate ln_bwd pass (read dz, x, dy; write dx) disappears.
template <int KS, int MODE, int TERMS, bool LN, bool TMA, int ABF = 0, int OBF = 0, bool LNB = false>
__global__ void __launch_bounds__(PM_THREADS + (TMA ? 32 : 0), 1)
    pm_gemm_kernel(const rcot_pm_params p, const PmGeom g, const __grid_constant__ CUtensorMap tm1,
                   const __grid_constant__ CUtensorMap tm2) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[PM_MAX_STAGES], empty_bar[PM_MAX_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint64_t raw_full[PM_RAW], raw_empty[PM_RAW];
  __shared__ uint32_t tmem_base_s;
  // LNB: per epilogue warp a 32 x 33 transpose pad (column sums over the warp's 32 pixels) and 2 x 256 running sums
  __shared__ float lnb_scr[LNB ? 4 * 32 * 33 : 1];
  __shared__ float lnb_acc[LNB ? 4 * 512 : 1];
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  constexpr int TAA = ABF ? 1 : TA;        // operand images of A: a bf16-stored A has no lo term
  static_assert(!LNB || (KS == 1 && !LN && !OBF), "LayerNorm-backward epilogue: 1x1 kernels with an fp32 output");
  if (LNB)
    for (int i = threadIdx.x; i < 4 * 512; i += blockDim.x) lnb_acc[i] = 0.f;
  static_assert(!(ABF || OBF) || (TMA && KS == 1 && !(ABF && LN)), "bf16 storage: TMA-staged 1x1 kernels only");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HWr = p.Hr * p.Wr, HWs = p.Hs * p.Ws;
  const int BN = g.BN, nk = g.nk, stages = g.stages;
  const uint32_t a_tile = op_tile_bytes(128);
  const uint32_t b_tile = op_tile_bytes(BN);          // this item's columns (a slice of the packed pass)
  const uint32_t b_tile_f = op_tile_bytes(g.BNf);     // the packed pass: [chunk][term][BNf x 32]
  const uint32_t stage_bytes = TAA * a_tile + TA * b_tile;
  const int total_tiles = g.tiles_m * g.passes;
  // Features of the conv instantiations only, resolved at compile time so that the 1x1 kernels (the hot path: the
  // epilogue of the write-heavy shapes lost 20 % when these were run-time branches) carry none of their code:
  const bool use_cls = (MODE == 1 && KS == 4) && g.cls;      // parity-class tiles (stride-2 transposed conv)
  const int nslice = (KS > 1) ? g.nslice : 1;                // N-sliced work items

  // LayerNorm affine parameters, interleaved (gamma, beta) and zero-padded to the chunk grid
  float* ln_gb = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);
  if (LN) {
    for (int k = tid; k < nk * KC; k += blockDim.x) {
      ln_gb[2 * k] = k < g.Ktot ? __ldg(p.ln_gamma + k) : 0.f;
      ln_gb[2 * k + 1] = k < g.Ktot ? __ldg(p.ln_beta + k) : 0.f;
    }
  }
  uint8_t* raw_ring = smem + (size_t)stages * stage_bytes + (LN ? (size_t)nk * KC * 2 * sizeof(float) : 0);
  if (warp == 0) tmem_alloc(&tmem_base_s, g.tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], PM_PROD_WARPS + 1);   // one arrival per producer warp + the TMA expect_tx
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);                  // one arrival per epilogue warp
    }
    if (TMA) {
      for (int i = 0; i < g.nraw; ++i) {
        mbar_init(&raw_full[i], 1);                 // the loader's expect_tx arrival (+ the copies' bytes)
        mbar_init(&raw_empty[i], PM_PROD_WARPS);    // one arrival per producer warp
      }
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (TMA && warp == PM_PROD_WARPS + 5) {
    // =========================================================== loader (TMA variant): raw fp32 tiles by bulk copy
    int rs = 0;
    uint32_t rph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int pass = t / g.tiles_m, mt = t - pass * g.tiles_m;
      int b, pix0;
      if (g.flat) {
        const long r = (long)mt * 128;
        b = (int)(r / HWr);
        pix0 = (int)(r - (long)b * HWr);
      } else {
        b = mt / g.tiles_per_img;
        pix0 = (mt - b * g.tiles_per_img) * 128;
      }
      for (int c = 0; c < nk; ++c) {
        mbar_wait(&raw_empty[rs], rph ^ 1);
        if (lane == 0) {
          // channels beyond the tensor (ragged last chunk) are filled with zeros by the copy engine and still
          // count towards the transaction bytes
          mbar_arrive_expect_tx(&raw_full[rs], ABF ? PM_RAW_BYTES / 2 : PM_RAW_BYTES);
          const int k = c * KC;
          const bool second = k >= p.C1;
          const CUtensorMap* tm = second ? &tm2 : &tm1;
          const int kc = second ? k - p.C1 : k;
          tensor_g2s_3d(raw_ring + (size_t)rs * PM_RAW_BYTES, tm, pix0, kc, b, &raw_full[rs]);
        }
        __syncwarp();
        if (++rs == g.nraw) {
          rs = 0;
          rph ^= 1;
        }
      }
    }
  } else if (TMA && warp < PM_PROD_WARPS) {
    // =========================================================== producers (TMA variant)
    // raw slot -> registers -> (LayerNorm) -> bf16 hi/lo operand stage.  No global loads except the per-tile
    // LayerNorm statistics; the weight stage is still one bulk copy issued by thread 0.
    const int row = tid & 127, khalf = tid >> 7;
    int rs = 0, ps_ = 0;
    uint32_t rph = 0, pph_ = 0;
    // (image, pixel) of this thread's row in tile t
    auto locate = [&](int t, int& b, int& pix) {
      const int pass = t / g.tiles_m, mt = t - pass * g.tiles_m;
      if (g.flat) {
        const long r = (long)mt * 128 + row;
        b = (int)(r / HWr);
        pix = (int)(r - (long)b * HWr);
      } else {
        b = mt / g.tiles_per_img;
        pix = (mt - b * g.tiles_per_img) * 128 + row;
      }
    };
    // LayerNorm statistics are requested one tile ahead (unconditionally, clamped to the last tile): at K = 48..96 a
    // tile is two or three chunks, and the exposed load was 43 % of the stall samples
    float2 st_next = make_float2(0.f, 0.f);
    if (LN && blockIdx.x < total_tiles) {
      int b0, p0;
      locate(blockIdx.x, b0, p0);
      st_next = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)b0 * HWs + p0);
    }
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int pass = t / g.tiles_m;
      int b, pix;
      locate(t, b, pix);
      float mu = 0.f, rstd = 0.f;
      if (LN) {
        if (p.debug & 1) {          // A/B knob: statistics loaded at the tile's start
          const float2 st = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)b * HWs + pix);
          mu = st.x;
          rstd = st.y;
        } else {
          mu = st_next.x;
          rstd = st_next.y;
          int bn, pn;
          locate(min(t + (int)gridDim.x, total_tiles - 1), bn, pn);
          st_next = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)bn * HWs + pn);
        }
      }
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (g.flat ? 0 : (size_t)b * p.wpack_bs) +
                            (size_t)pass * g.nk_full * (2 * b_tile_f);
      for (int c = 0; c < nk; ++c) {
        mbar_wait(&raw_full[rs], rph);
        const int k0 = c * KC + khalf * 16;
        float v[16];
        uint32_t pk[8];
        if (ABF) {
          // bf16 source: the 16 values ARE the operand (no split, no lo image); channels beyond the tensor were
          // zero-filled by the copy engine
          const unsigned short* raw =
              reinterpret_cast<const unsigned short*>(raw_ring + (size_t)rs * PM_RAW_BYTES) + (khalf * 16) * 128 + row;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t lo16 = (k0 + 2 * i < g.Ktot) ? raw[(2 * i) * 128] : 0u;
            const uint32_t hi16 = (k0 + 2 * i + 1 < g.Ktot) ? raw[(2 * i + 1) * 128] : 0u;
            pk[i] = lo16 | (hi16 << 16);
          }
        } else {
          const float* raw = reinterpret_cast<const float*>(raw_ring + (size_t)rs * PM_RAW_BYTES) + (khalf * 16) * 128 + row;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (k0 + i < g.Ktot) ? raw[i * 128] : 0.f;   // never-copied rows read as 0
        }
        const int s = ps_;
        const uint32_t ph = pph_;
        if (++ps_ == stages) {
          ps_ = 0;
          pph_ ^= 1;
        }
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        if (tid == 0) {
          mbar_arrive_expect_tx(&full_bar[s], TA * b_tile);
          bulk_g2s(st + TAA * a_tile, wsrc + (size_t)c * (2 * b_tile_f), TA * b_tile, &full_bar[s]);
        }
        if (ABF) {
          const uint32_t pa[4] = {pk[0], pk[1], pk[2], pk[3]}, pb[4] = {pk[4], pk[5], pk[6], pk[7]};
          op_store8_bf16(st, row, khalf * 2, pa);
          op_store8_bf16(st, row, khalf * 2 + 1, pb);
        } else {
          if (LN) {
            const float* gb = ln_gb + k0 * 2;   // interleaved (gamma, beta), zero beyond Ktot
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 w2 = *reinterpret_cast<const float2*>(gb + 2 * i);
              v[i] = (v[i] - mu) * rstd * w2.x + w2.y;
            }
          }
          op_store8<TERMS>(st, st + a_tile, row, khalf * 2, v);
          op_store8<TERMS>(st, st + a_tile, row, khalf * 2 + 1, v + 8);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&full_bar[s]);
          mbar_arrive(&raw_empty[rs]);       // every lane's raw values have been consumed (stored) by now
        }
        if (++rs == g.nraw) {
          rs = 0;
          rph ^= 1;
        }
      }
    }
  } else if (!TMA && warp < PM_PROD_WARPS) {
    // =========================================================== producers
    // Work items are (tile, K-chunk) pairs; the 16 loads of item j+1 are issued before item j is
    // converted, so global-load latency overlaps the conversion and the barrier waits.
    const int row = tid & 127, khalf = tid >> 7;  // this thread fills k8 groups {2*khalf, 2*khalf+1}
    struct TileCtx {
      const float *b1, *b2;   // image base pointers (+ pixel offset for the 1x1 fast path)
      const uint8_t* wsrc;
      float mu, rstd;
      int ry, rx;
      int cls;                // parity class (y&1)*2 + (x&1) of the tile's pixels (g.cls only)
      bool valid;
    };
    auto setup = [&](int t) {
      TileCtx x;
      const int pass = t / g.tiles_m, mt = t - pass * g.tiles_m;
      int b, pix;
      x.cls = 0;
      if (use_cls) {
        const int cl = mt / g.tpc;
        const long r = (long)(mt - cl * g.tpc) * 128 + row;
        const int q = HWr >> 2, hw2 = p.Wr >> 1;          // pixels per class and image, class-grid width
        x.valid = r < (long)q * p.B;
        b = x.valid ? (int)(r / q) : 0;
        const int rem = x.valid ? (int)(r - (long)b * q) : 0;
        const int yy = rem / hw2, xx = rem - yy * hw2;
        pix = (2 * yy + (cl >> 1)) * p.Wr + 2 * xx + (cl & 1);
        x.cls = cl;
      } else if (g.flat) {
        const long r = (long)mt * 128 + row;
        x.valid = r < g.rows_total;
        b = x.valid ? (int)(r / HWr) : 0;
        pix = x.valid ? (int)(r - (long)b * HWr) : 0;
      } else {
        b = mt / g.tiles_per_img;
        pix = (mt - b * g.tiles_per_img) * 128 + row;
        x.valid = pix < HWr;
        if (!x.valid) pix = 0;
      }
      {
        const int ry = pix / p.Wr, rx = pix - ry * p.Wr;
        x.ry = (MODE == 0) ? ry * p.stride - p.pad : ry + p.pad;
        x.rx = (MODE == 0) ? rx * p.stride - p.pad : rx + p.pad;
      }
      const int poff = (KS == 1) ? pix : 0;
      x.b1 = p.in + (size_t)b * p.in_bs + poff;
      x.b2 = p.in2 ? p.in2 + (size_t)b * p.in2_bs + poff : nullptr;
      x.mu = 0.f;
      x.rstd = 0.f;
      if (LN && x.valid) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + (size_t)b * HWs + pix);
        x.mu = st.x;
        x.rstd = st.y;
      }
      x.wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (g.flat ? 0 : (size_t)b * p.wpack_bs) +
               (size_t)(pass / nslice) * g.nk_full * (2 * b_tile_f) + (size_t)(pass % nslice) * b_tile;
      return x;
    };
    auto load16 = [&](float* v, const TileCtx& x, int c) {
      const int k0 = c * KC + khalf * 16;   // this thread's 16 consecutive K indices of the chunk
      if (KS == 1) {
        if (x.valid && g.c1_aligned && k0 + 16 <= g.Ktot) {
          // fast path: 16 channels of ONE source tensor at this pixel -> base + i*HW, no predicates
          const float* sp = (k0 < p.C1) ? x.b1 + (size_t)k0 * HWs : x.b2 + (size_t)(k0 - p.C1) * HWs;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __ldg(sp + (size_t)i * HWs);
          return;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k = k0 + i;
          float val = 0.f;
          if (x.valid && k < g.Ktot)
            val = __ldg((k < p.C1) ? (x.b1 + (size_t)k * HWs) : (x.b2 + (size_t)(k - p.C1) * HWs));
          v[i] = val;
        }
        return;
      }
      if (p.tap_major) {
        // K = (ky, kx, channel), C1 % 16 == 0: this thread's 16 consecutive K indices are 16 channels at ONE tap
        // -> one bounds test, strided reads (the K padding of the last chunk maps to tap >= KS*KS: zeros)
        const int k0 = c * KC + khalf * 16;
        int tap = k0 / p.C1;
        const int ch0 = k0 - tap * p.C1;
        if (use_cls) {
          // c runs over this class's 4 taps: (ky, kx) = (ky0 + 2*(tap>>1), kx0 + 2*(tap&1))
          const int ky0 = ((x.cls >> 1) + p.pad) & 1, kx0 = ((x.cls & 1) + p.pad) & 1;
          tap = (ky0 + 2 * (tap >> 1)) * KS + kx0 + 2 * (tap & 1);
        }
        const int ky = tap / KS, kx = tap - ky * KS;
        int sy, sx;
        bool ok = x.valid && tap < KS * KS;
        if (MODE == 0) {
          sy = x.ry + ky;
          sx = x.rx + kx;
        } else {
          const int ty = x.ry - ky, tx = x.rx - kx;
          if (p.stride == 1) {
            sy = ty;
            sx = tx;
          } else {
            sy = ty >> 1;
            sx = tx >> 1;
            ok = ok && (((ty | tx) & 1) == 0);
          }
        }
        ok = ok && ((unsigned)sy < (unsigned)p.Hs) && ((unsigned)sx < (unsigned)p.Ws);
        if (ok) {
          const float* sp = x.b1 + (size_t)ch0 * HWs + sy * p.Ws + sx;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __ldg(sp + (size_t)i * HWs);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        return;
      }
      // general conv geometry: decompose k0 once, then step (ch, ky, kx)
      constexpr int KK = KS * KS;
      int ch = k0 / KK;
      const int r0 = k0 - ch * KK;
      int ky = r0 / KS, kx = r0 - ky * KS;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float val = 0.f;
        if (x.valid && k0 + i < g.Ktot) {
          int sy, sx;
          bool ok;
          if (MODE == 0) {
            sy = x.ry + ky;     // ry pre-scaled: ry*stride - pad
            sx = x.rx + kx;
            ok = true;
          } else {
            const int ty = x.ry - ky, tx = x.rx - kx;   // ry pre-offset: ry + pad
            if (p.stride == 1) {
              sy = ty;
              sx = tx;
              ok = true;
            } else {            // stride 2 (checked by the launcher)
              sy = ty >> 1;
              sx = tx >> 1;
              ok = ((ty | tx) & 1) == 0;
            }
          }
          ok = ok && ((unsigned)sy < (unsigned)p.Hs) && ((unsigned)sx < (unsigned)p.Ws);
          if (ok) {
            const float* sp = (ch < p.C1) ? (x.b1 + (size_t)ch * HWs) : (x.b2 + (size_t)(ch - p.C1) * HWs);
            val = __ldg(sp + sy * p.Ws + sx);
          }
        }
        v[i] = val;
        if (++kx == KS) {
          kx = 0;
          if (++ky == KS) {
            ky = 0;
            ++ch;
          }
        }
      }
    };
    int ps_ = 0;
    uint32_t pph_ = 0;
    // The loads of items j+1 and j+2 are in flight while item j is converted (two-deep register prefetch:
    // 2 x 16 loads x 256 threads = 32 KB outstanding per SM).
    int t = blockIdx.x, c = 0;                          // walker: the next item to issue loads for
    TileCtx xw = setup(t < total_tiles ? t : 0);
    auto step_walker = [&]() {
      if (++c == nk) {
        c = 0;
        t += gridDim.x;
        if (t < total_tiles) xw = setup(t);
      }
    };
    struct Slot {          // one prefetched work item: 16 operand values + where they belong
      float v[16];
      TileCtx x;
      int c;
      bool ok;
    };
    auto fetch = [&](Slot& S) {
      S.ok = t < total_tiles;
      if (S.ok) {
        S.x = xw;
        S.c = c;
        load16(S.v, xw, c);
        step_walker();
      }
    };
    auto process = [&](Slot& S) {   // convert + store one item into the next pipeline stage
      const int s = ps_;
      const uint32_t ph = pph_;
      if (++ps_ == stages) {
        ps_ = 0;
        pph_ ^= 1;
      }
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* st = smem + (size_t)s * stage_bytes;
      if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar[s], TA * b_tile);
        int wc = S.c;               // chunk of the packed weights (tap-major: chunk = tap * C1/32 + channel chunk)
        if (use_cls) {
          const int cpt = p.C1 >> 5, te = S.c / cpt;
          const int ky0 = ((S.x.cls >> 1) + p.pad) & 1, kx0 = ((S.x.cls & 1) + p.pad) & 1;
          wc = ((ky0 + 2 * (te >> 1)) * KS + kx0 + 2 * (te & 1)) * cpt + (S.c - te * cpt);
        }
        const uint8_t* wsrc = S.x.wsrc + (size_t)wc * (2 * b_tile_f);
        if (nslice == 1) {
          bulk_g2s(st + TAA * a_tile, wsrc, TA * b_tile, &full_bar[s]);          // hi and lo images are adjacent
        } else {
          bulk_g2s(st + TAA * a_tile, wsrc, b_tile, &full_bar[s]);
          if (TA > 1) bulk_g2s(st + TAA * a_tile + b_tile, wsrc + b_tile_f, b_tile, &full_bar[s]);
        }
      }
      if (LN) {
        const float* gb = ln_gb + (S.c * KC + khalf * 16) * 2;   // interleaved (gamma, beta), zero beyond Ktot
        if (S.x.valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 w2 = *reinterpret_cast<const float2*>(gb + 2 * i);
            S.v[i] = (S.v[i] - S.x.mu) * S.x.rstd * w2.x + w2.y;
          }
        }
      }
      op_store8<TERMS>(st, st + a_tile, row, khalf * 2, S.v);
      op_store8<TERMS>(st, st + a_tile, row, khalf * 2 + 1, S.v + 8);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    };
    // three register sets rotate (no copies): while one item is converted, the loads of the next two fly
    Slot A, B, C;
    fetch(A);
    fetch(B);
    while (true) {
      fetch(C);
      if (!A.ok) break;
      process(A);
      fetch(A);
      if (!B.ok) break;
      process(B);
      fetch(B);
      if (!C.ok) break;
      process(C);
    }
  } else if (warp == PM_PROD_WARPS) {
    // =========================================================== MMA issuer (whole warp waits, lane 0 issues)
    const uint32_t idesc = make_idesc_bf16(128, BN);
    uint32_t tcount = 0;
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1, aph = (tcount >> 1) & 1;
      mbar_wait(&acc_empty[buf], aph ^ 1);
      tc_fence_after();
      const uint32_t d = tmem + buf * BN;
      for (int c = 0; c < nk; ++c) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
          issue_stage<TERMS, !ABF>(d, st, st + a_tile, st + TAA * a_tile, st + TAA * a_tile + b_tile, idesc, c == 0);
          tc_commit(&empty_bar[s]);
        }
        __syncwarp();
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (elect_one()) tc_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else if (warp < PM_PROD_WARPS + 5) {
    // =========================================================== epilogue (warp % 4 = TMEM lane quarter)
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    // fast epilogues (operand of the fused op prefetched one 16-channel group ahead of the stores): plain, residual
    // add, and the LeakyReLU-derivative mask of F_net's data gradients; everything else takes the generic path
    const bool epi_mask = (KS > 1) && p.mask_y && !p.bias && !p.act && !p.accumulate && !p.residual;
    const bool epi_other = (p.bias || p.act || p.mask_y || p.accumulate) && !epi_mask;
    const bool epi_plain = !epi_other && !epi_mask && !p.residual, epi_res = !epi_other && !epi_mask && p.residual;
    uint32_t tcount = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tcount) {
      const int pass = t / g.tiles_m, mt = t - pass * g.tiles_m;
      int b, pix;
      bool valid;
      if (use_cls) {
        const int cl = mt / g.tpc;
        const long r = (long)(mt - cl * g.tpc) * 128 + row;
        const int q = HWr >> 2, hw2 = p.Wr >> 1;
        valid = r < (long)q * p.B;
        b = valid ? (int)(r / q) : 0;
        const int rem = valid ? (int)(r - (long)b * q) : 0;
        const int yy = rem / hw2, xx = rem - yy * hw2;
        pix = (2 * yy + (cl >> 1)) * p.Wr + 2 * xx + (cl & 1);
      } else if (g.flat) {
        const long r = (long)mt * 128 + row;
        valid = r < g.rows_total;
        b = valid ? (int)(r / HWr) : 0;
        pix = valid ? (int)(r - (long)b * HWr) : 0;
      } else {
        b = mt / g.tiles_per_img;
        pix = (mt - b * g.tiles_per_img) * 128 + row;
        valid = pix < HWr;
      }
      if (TMA && epi_res && !(p.debug & 2)) {
        // pull the residual rows of this CTA's NEXT tile into L2 while this tile is drained (TMA variant: whole
        // 128-pixel tiles inside one image): the residual loads below run only 16 channels ahead of the stores
        const int tn = t + gridDim.x;
        if (tn < total_tiles) {
          const int passn = tn / g.tiles_m, mtn = tn - passn * g.tiles_m;
          int bn, pixn;
          if (g.flat) {
            const long r = (long)mtn * 128;
            bn = (int)(r / HWr);
            pixn = (int)(r - (long)bn * HWr);
          } else {
            bn = mtn / g.tiles_per_img;
            pixn = (mtn - bn * g.tiles_per_img) * 128;
          }
          const int nb0 = passn * BN;
          const int nc = max(0, min(BN, p.N - nb0));
          const float* rn = p.residual + (size_t)bn * p.res_bs + (size_t)nb0 * HWr + pixn;
          const int et = (warp & 3) * 32 + lane;                       // 0..127 over the four epilogue warps
          for (int L = et; L < 4 * nc; L += 128)                       // 128-byte lines: channel L/4, quarter L%4
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rn + (size_t)(L >> 2) * HWr + (L & 3) * 32));
        }
      }
      const uint32_t buf = tcount & 1, aph = (tcount >> 1) & 1;
      mbar_wait(&acc_full[buf], aph);
      tc_fence_after();
      if (LNB) {
        // ---------------- LayerNorm backward on the accumulator row (single pass: nbase = 0, all N = C columns)
        const int Cn = p.N, ng = (Cn + 15) >> 4;
        const uint32_t tb = lane_base + buf * BN;
        const size_t po = (size_t)b * HWr + pix;
        float mu = 0.f, rstd = 0.f;
        if (valid) {
          const float2 st2 = __ldg(reinterpret_cast<const float2*>(p.lnb_stats) + po);
          mu = st2.x;
          rstd = st2.y;
        }
        const float* xr = p.lnb_x + (size_t)b * p.lnb_x_bs + pix;
        const float* dyr = p.residual ? p.residual + (size_t)b * p.res_bs + pix : nullptr;
        float* orow = p.out + (size_t)b * p.out_bs + pix;
        float* scr = lnb_scr + (warp & 3) * (32 * 33);
        float* accw = lnb_acc + (warp & 3) * 512;
        float s1 = 0.f, s2 = 0.f;
        for (int gi = 0; gi < ng; ++gi) {
          float dz[16], xv[16];
          tmem_ld16(tb + gi * 16, dz);
          const int nrem = Cn - gi * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) xv[i] = (valid && i < nrem) ? __ldg(xr + (size_t)(gi * 16 + i) * HWr) : 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool ok = valid && i < nrem;
            const float xh = (xv[i] - mu) * rstd;
            const float d = ok ? dz[i] : 0.f;
            const float gq = d * (i < nrem ? __ldg(p.lnb_gamma + gi * 16 + i) : 0.f);
            s1 += gq;
            s2 = fmaf(gq, xh, s2);
            scr[lane * 33 + i] = d * xh;          // -> dgamma
            scr[lane * 33 + 16 + i] = d;          // -> dbeta
          }
          __syncwarp();
          float tcol = 0.f;
#pragma unroll 8
          for (int l = 0; l < 32; ++l) tcol += scr[l * 33 + lane];
          accw[gi * 32 + lane] += tcol;           // lane < 16: dgamma of channel gi*16+lane; else dbeta of gi*16+lane-16
          __syncwarp();
        }
        const float invC = 1.f / (float)Cn;
        const float m1 = s1 * invC, m2 = s2 * invC;
        for (int gi = 0; gi < ng; ++gi) {
          float dz[16], xv[16], dyv[16];
          tmem_ld16(tb + gi * 16, dz);
          const int nrem = Cn - gi * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool ok = valid && i < nrem;
            xv[i] = ok ? __ldg(xr + (size_t)(gi * 16 + i) * HWr) : 0.f;
            dyv[i] = (ok && dyr) ? __ldg(dyr + (size_t)(gi * 16 + i) * HWr) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (valid && i < nrem) {
              const float xh = (xv[i] - mu) * rstd;
              const float gq = dz[i] * __ldg(p.lnb_gamma + gi * 16 + i);
              orow[(size_t)(gi * 16 + i) * HWr] = dyv[i] + rstd * (gq - m1 - xh * m2);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      const int nbase = (pass / nslice) * g.BNf + (pass % nslice) * BN;
      float* o = p.out + (size_t)b * p.out_bs + (size_t)(p.out_coff + nbase) * HWr + pix;
      const float* mk = p.mask_y ? p.mask_y + (size_t)b * p.mask_bs + (size_t)nbase * HWr + pix : nullptr;
      const float* rs = p.residual ? p.residual + (size_t)b * p.res_bs + (size_t)nbase * HWr + pix : nullptr;
      const bool st_ok = valid;
      const int ncols = max(0, min(BN, p.N - nbase));    // valid columns of this item
      const int ngroups = (ncols + 15) >> 4;             // 16-column groups
      uint32_t ra[16], rb[16];
      float qa[16], qb[16];                              // residual values, fetched one group ahead
      float ln_s0 = 0.f, ln_s1 = 0.f, ln_s2 = 0.f;       // shifted sums for the optional LayerNorm statistics
      const bool want_stats = p.stats_out != nullptr;
      const float* pre = epi_mask ? mk : rs;   // tensor whose values are fetched ahead (same indexing as out)
      auto ldres = [&](float (&q)[16], int gi) {
        const int nrem = ncols - gi * 16;
        const float* r0 = pre + (size_t)(gi * 16) * HWr;
#pragma unroll
        for (int i = 0; i < 16; ++i) q[i] = (st_ok && i < nrem) ? __ldg(r0 + (size_t)i * HWr) : 0.f;
      };
      auto emit = [&](const uint32_t (&cur)[16], const float (&q)[16], int gi) {
        const int nrem = ncols - gi * 16;
        if (!st_ok) return;
        float* og = o + (size_t)(gi * 16) * HWr;
        if (epi_plain || epi_res || epi_mask) {
          float y[16];
          if (epi_mask) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(cur[i]) * (q[i] > 0.f ? 1.f : p.slope);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(cur[i]) + (epi_res ? q[i] : 0.f);
          }
          if (OBF) {
            // bf16 output tensor (same element offsets), 2-byte stores: a warp writes 64 contiguous bytes.  (Trading
            // values between neighbouring lanes to store pixel pairs as 4 bytes was measured slower: 0.54 vs 0.36 ms for
            // x -> u at C=96, 128x128, B=32 -- the shuffles cost more issue slots than the narrower stores.)
            __nv_bfloat16* og16 = reinterpret_cast<__nv_bfloat16*>(p.out) + (og - p.out);
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nrem) og16[(size_t)i * HWr] = __float2bfloat16_rn(y[i]);
          } else if (nrem >= 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) og[(size_t)i * HWr] = y[i];
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nrem) og[(size_t)i * HWr] = y[i];
          }
          if (want_stats) {
            if (gi == 0) ln_s0 = y[0];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float d = (i < nrem) ? y[i] - ln_s0 : 0.f;
              ln_s1 += d;
              ln_s2 = fmaf(d, d, ln_s2);
            }
          }
        } else {
          const float* mg = mk ? mk + (size_t)(gi * 16) * HWr : nullptr;
          const float* rg = rs ? rs + (size_t)(gi * 16) * HWr : nullptr;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < nrem) {
              float y = __uint_as_float(cur[i]);
              if (p.bias) y += __ldg(p.bias + nbase + gi * 16 + i);
              if (p.act) y = y > 0.f ? y : y * p.slope;
              if (mg) y *= (__ldg(mg + (size_t)i * HWr) > 0.f) ? 1.f : p.slope;
              if (rg) y += __ldg(rg + (size_t)i * HWr);
              if (p.accumulate) y += og[(size_t)i * HWr];
              og[(size_t)i * HWr] = y;
            }
          }
        }
      };
      const uint32_t tbase = lane_base + buf * BN;
      const bool pref = epi_res || epi_mask;
      tmem_ld16_nowait(tbase, ra);
      if (pref) ldres(qa, 0);
      for (int gi = 0; gi < ngroups; gi += 2) {   // TMEM + residual loads run one group ahead of the stores
        tmem_ld_wait();
        if (gi + 1 < ngroups) {
          tmem_ld16_nowait(tbase + (gi + 1) * 16, rb);
          if (pref) ldres(qb, gi + 1);
        }
        emit(ra, qa, gi);
        if (gi + 1 < ngroups) {
          tmem_ld_wait();
          if (gi + 2 < ngroups) {
            tmem_ld16_nowait(tbase + (gi + 2) * 16, ra);
            if (pref) ldres(qa, gi + 2);
          }
          emit(rb, qb, gi + 1);
        }
      }
      if (want_stats && st_ok) {
        const float inv = 1.f / (float)p.N;
        const float m = ln_s1 * inv;
        const float var = fmaxf(ln_s2 * inv - m * m, 0.f);
        reinterpret_cast<float2*>(p.stats_out)[(size_t)b * HWr + pix] = make_float2(ln_s0 + m, 1.0f / sqrtf(var + 1e-5f));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  if (LNB && warp >= PM_PROD_WARPS + 1 && warp < PM_PROD_WARPS + 5) {
    // flush this warp's running column sums (every lane only ever touched its own slots)
    const float* accw = lnb_acc + (warp & 3) * 512;
    const int ng = (p.N + 15) >> 4;
    for (int gi = 0; gi < ng; ++gi) {
      const int ch = gi * 16 + (lane & 15);
      if (ch < p.N) atomicAdd((lane < 16 ? p.lnb_dgamma : p.lnb_dbeta) + ch, accw[gi * 32 + lane]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, g.tmem_cols);
}

static int g_num_sms = 0;

template <int KS, int MODE, int TERMS, bool LN, bool TMA = false, int ABF = 0, int OBF = 0, bool LNB = false>
static int launch_pm(const rcot_pm_params& p, cudaStream_t stream) {
  PmGeom g;
  g.Ktot = (p.C1 + p.C2) * KS * KS;
  g.nk = cdiv(g.Ktot, KC);
  NPlan pl = make_nplan(p.N);
  g.BN = pl.BN;
  g.BNf = pl.BN;
  g.nslice = 1;
  g.passes = pl.passes;
  {
    // few M tiles and a long K loop (conv with K >= 1024): slice the N pass while that still fills idle SMs
    const long HW0 = (long)p.Hr * p.Wr;
    const long tm = (p.wpack_bs == 0) ? cdiv(HW0 * p.B, 128) : (long)cdiv(HW0, 128) * p.B;
    const int ktot = (p.C1 + p.C2) * KS * KS;
    while (KS > 1 && ktot >= 1024 && !p.stats_out && g.nslice < 4 && tm * pl.passes * g.nslice * 2 <= 148 &&
           (g.BN / 2) % 16 == 0) {
      g.nslice *= 2;
      g.BN /= 2;
    }
    g.passes = pl.passes * g.nslice;
  }
  constexpr int TA = (TERMS > 1) ? 2 : 1;
  constexpr int TAA = ABF ? 1 : TA;
  const size_t stage_bytes = (size_t)TAA * op_tile_bytes(128) + (size_t)TA * op_tile_bytes(g.BN);
  const size_t ln_bytes = LN ? (size_t)g.nk * KC * 2 * sizeof(float) : 0;
  g.c1_aligned = (p.C2 == 0 || p.C1 % 16 == 0) ? 1 : 0;
  static int nraw_env = -1;           // A/B switch RCOT_PM_RAW: slots of the raw fp32 ring (16 KB each; default 4)
  if (nraw_env < 0) {
    const char* e = getenv("RCOT_PM_RAW");
    nraw_env = e ? atoi(e) : 4;
    if (nraw_env < 2) nraw_env = 2;
    if (nraw_env > PM_RAW) nraw_env = PM_RAW;
  }
  g.nraw = nraw_env;
  const size_t raw_bytes = TMA ? (size_t)g.nraw * PM_RAW_BYTES : 0;
  const size_t lnb_bytes = LNB ? (4 * 32 * 33 + 4 * 512) * sizeof(float) : 0;    // static shared memory of the LNB epilogue
  int stages = (int)((196 * 1024 - raw_bytes - ln_bytes - lnb_bytes) / stage_bytes);
  if (stages > PM_MAX_STAGES) stages = PM_MAX_STAGES;
  if (stages < 2) stages = 2;
  g.stages = stages;
  const long HWr = (long)p.Hr * p.Wr;
  g.flat = (p.wpack_bs == 0) ? 1 : 0;
  g.rows_total = HWr * p.B;
  g.tiles_per_img = cdiv(HWr, 128);
  g.tiles_m = g.flat ? cdiv(g.rows_total, 128) : g.tiles_per_img * p.B;
  g.tmem_cols = tmem_cols_pow2(2 * g.BN);
  g.nk_full = g.nk;
  g.cls = 0;
  g.tpc = 0;
  if (MODE == 1 && KS == 4 && p.stride == 2 && p.tap_major && p.C1 % 32 == 0 && g.flat && p.Hr % 2 == 0 &&
      p.Wr % 2 == 0) {
    g.cls = 1;
    g.tpc = cdiv(HWr / 4 * p.B, 128);
    g.tiles_m = 4 * g.tpc;
    g.nk = 4 * (p.C1 / 32);
  }
  const size_t smem = stages * stage_bytes + ln_bytes + raw_bytes;
  RCOT_REQUIRE(smem + lnb_bytes <= 208 * 1024, "pm_gemm: %zu bytes of shared memory needed", smem + lnb_bytes);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pm_gemm_kernel<KS, MODE, TERMS, LN, TMA, ABF, OBF, LNB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(208 * 1024 - lnb_bytes));
    if (e != cudaSuccess) {
      set_error("pm_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const long total = (long)g.tiles_m * g.passes;
  const int grid = (int)(total < g_num_sms ? total : g_num_sms);
  CUtensorMap tm1, tm2;
  memset(&tm1, 0, sizeof(tm1));
  memset(&tm2, 0, sizeof(tm2));
  if (TMA) {
    int rc = make_act_map(&tm1, p.in, p.in_bs, p.C1, HWr, p.B, 128, KC, "pm_gemm", ABF);
    if (rc == RCOT_OK && p.in2) rc = make_act_map(&tm2, p.in2, p.in2_bs, p.C2, HWr, p.B, 128, KC, "pm_gemm");
    if (rc != RCOT_OK) return rc;
  }
  pm_gemm_kernel<KS, MODE, TERMS, LN, TMA, ABF, OBF, LNB><<<grid, PM_THREADS + (TMA ? 32 : 0), smem, stream>>>(p, g, tm1, tm2);
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
  RCOT_REQUIRE(!(p.in_bf16 || p.out_bf16) || p.ks == 1, "pm_gemm: bf16 tensors are supported by the 1x1 kernels only");
  RCOT_REQUIRE(p.mode == 0 || p.mode == 1, "pm_gemm: mode must be 0 or 1");
  const bool ln = p.ln_stats != nullptr;
  if (ln) RCOT_REQUIRE(p.ks == 1 && p.ln_gamma && p.ln_beta, "pm_gemm: LayerNorm prologue needs ks==1, gamma, beta");
  if (p.ks == 1)
    RCOT_REQUIRE(p.stride == 1 && p.pad == 0 && p.Hs == p.Hr && p.Ws == p.Wr, "pm_gemm: 1x1 needs stride 1, pad 0");
  RCOT_REQUIRE(p.stride == 1 || p.stride == 2, "pm_gemm: stride must be 1 or 2");
  if (p.stats_out)
    RCOT_REQUIRE(p.N <= 256 && !p.bias && !p.act && !p.mask_y && !p.accumulate && p.out_coff == 0,
                 "pm_gemm: stats_out needs N <= 256 and the plain / residual epilogue");
  if (p.tap_major)
    RCOT_REQUIRE(p.ks > 1 && p.C2 == 0 && p.C1 % 16 == 0, "pm_gemm: tap_major needs ks > 1, no concat, C1 %% 16 == 0");
#define PM_DISPATCH(KS, MODE, LN)                                           \
  return (p.terms == 3) ? launch_pm<KS, MODE, 3, LN>(p, stream) : launch_pm<KS, MODE, 1, LN>(p, stream)
  switch (p.ks) {
    case 1: {
      // TMA-staged A operand: whole 128-pixel tiles inside one image, 16-byte aligned channel rows, K chunks that
      // do not straddle the two concat sources.  RCOT_PM_TMA=0 switches back to the register-prefetch producers.
      static int tma_on = -1;
      if (tma_on < 0) {
        const char* e = getenv("RCOT_PM_TMA");
        tma_on = (e && e[0] == '0') ? 0 : 1;
      }
      const long HW = (long)p.Hr * p.Wr;
      const bool tma = tma_on && tensor_map_encoder() != nullptr && HW % 128 == 0 && p.in_bs % 4 == 0 && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 &&
                       (p.in2 == nullptr || (p.C1 % 32 == 0 && p.in2_bs % 4 == 0 &&
                                             (reinterpret_cast<uintptr_t>(p.in2) & 15) == 0));
      if (p.lnb_x) {
        // LayerNorm-backward epilogue (the dz GEMM of a block's backward)
        RCOT_REQUIRE(p.lnb_stats && p.lnb_gamma && p.lnb_dgamma && p.lnb_dbeta, "pm_gemm: LayerNorm-backward epilogue needs "
                     "stats, gamma, dgamma and dbeta");
        RCOT_REQUIRE(p.N <= 256 && !ln && !p.out_bf16 && !p.bias && !p.act && !p.mask_y && !p.accumulate && !p.stats_out &&
                     p.out_coff == 0 && p.in2 == nullptr && p.wpack_bs == 0,
                     "pm_gemm: LayerNorm-backward epilogue needs N <= 256 and the plain / residual epilogue");
        RCOT_REQUIRE(!p.in_bf16 || tma, "pm_gemm: bf16 input needs the TMA-staged path");
#define PM_LNB(T, A) \
  return (p.terms == 3) ? launch_pm<1, 0, 3, false, T, A, 0, true>(p, stream) : launch_pm<1, 0, 1, false, T, A, 0, true>(p, stream)
        if (tma && p.in_bf16) { PM_LNB(true, 1); }
        if (tma) { PM_LNB(true, 0); }
        PM_LNB(false, 0);
#undef PM_LNB
      }
      if (p.in_bf16 || p.out_bf16) {
        // bf16-storage mode of the hidden tensors: TMA-staged kernels only
        RCOT_REQUIRE(tma && p.in2 == nullptr, "pm_gemm: bf16 tensors need the TMA-staged 1x1 path (H*W %% 128 == 0, 16-byte "
                     "aligned, no concat); got %dx%d", p.Hr, p.Wr);
        RCOT_REQUIRE(!(p.in_bf16 && ln), "pm_gemm: LayerNorm prologue takes an fp32 input");
        if (p.out_bf16)
          RCOT_REQUIRE(!p.bias && !p.act && !p.mask_y && !p.accumulate && !p.residual && !p.stats_out,
                       "pm_gemm: a bf16 output takes the plain epilogue only");
#define PM_BF(LNF, A, O) \
  return (p.terms == 3) ? launch_pm<1, 0, 3, LNF, true, A, O>(p, stream) : launch_pm<1, 0, 1, LNF, true, A, O>(p, stream)
        if (p.in_bf16 && p.out_bf16) { PM_BF(false, 1, 1); }
        if (p.in_bf16) { PM_BF(false, 1, 0); }
        if (ln) { PM_BF(true, 0, 1); }
        PM_BF(false, 0, 1);
#undef PM_BF
      }
      if (tma) {
        if (ln) return (p.terms == 3) ? launch_pm<1, 0, 3, true, true>(p, stream) : launch_pm<1, 0, 1, true, true>(p, stream);
        return (p.terms == 3) ? launch_pm<1, 0, 3, false, true>(p, stream) : launch_pm<1, 0, 1, false, true>(p, stream);
      }
      if (ln) { PM_DISPATCH(1, 0, true); }
      PM_DISPATCH(1, 0, false);
    }
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
