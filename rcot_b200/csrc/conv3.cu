// conv3.cu -- direct CUDA-core kernels for the convolutions with THREE channels on one side
// (Net_Restormer.py:117 patch_embed 3->48, :326 output 96->3, :443 F_net features.0 3->64 k5, and their gradients).
// As implicit GEMMs on tcgen05 these are the worst launches of a training step: N = 3 is padded to 16 accumulator
// columns while the producers still gather Cin*k*k im2col columns per pixel tile (the 64 -> 3, 5x5 data gradient of
// F_net's first layer took 1.02 ms, the 96 -> 3 output conv 0.54 ms at 128x128, batch 32: 4-7 % of either roofline).
// The arithmetic is small (<= 5 GFLOP), so a register-tiled FP32 kernel is the right tool:
//
//   conv_to3: out[b, c, y, x] = sum_{ci, ky, kx} in[b, ci, y + ky - p, x + kx - p] * Wf[c, ci, ky, kx],  c < 3
//     forward  (output conv):            Wf[c, ci, ky, kx] = weight[c, ci, ky, kx]
//     data gradient of a 3 -> Cout conv: Wf[c, o, ky, kx]  = weight[o, c, k-1-ky, k-1-kx]      (stride 1, pad (k-1)/2)
//   A CTA owns a 64 x 16-pixel tile of one image, a thread a 1 x 4 strip and all three outputs; the input tile (+halo)
//   goes through shared memory in chunks of 8 channels, the whole (re-laid-out) weight tensor sits in shared memory:
//   per (channel, ky) a thread issues 2 + 4 16-byte shared loads for 60 FMAs (k = 5; 2 + 3 for 36 at k = 3).
#include "../../include/rcot_b200.h"
#include "common.cuh"

namespace rcot {

constexpr int C3_TW = 64, C3_TH = 16;   // pixel tile
constexpr int C3_CC = 8;                // channels per shared-memory chunk

template <int KS>
struct C3Geom {
  static constexpr int P = (KS - 1) / 2;
  static constexpr int ROWS = C3_TH + KS - 1;
  static constexpr int TWP = 72;                          // row stride (floats): >= 64 + KS - 1, multiple of 4
  static constexpr int WROW = (3 * KS + 3) / 4 * 4;       // taps of one (channel, ky): [kx][c], padded to 16 bytes
  static constexpr int CHUNK = C3_CC * ROWS * TWP;        // floats of one data chunk
};

template <int KS>
__global__ void __launch_bounds__(256) conv_to3_kernel(const float* __restrict__ in, int64_t in_bs, const float* __restrict__ w,
                                                       int w_sc, int w_sci, int flip, float* __restrict__ out, int64_t out_bs,
                                                       const float* __restrict__ res, int64_t res_bs, int Cin, int H, int W) {
  using G = C3Geom<KS>;
  extern __shared__ __align__(16) float c3sm[];
  float* wS = c3sm;                                   // [Cin][KS][WROW]
  float* dS = c3sm + (size_t)Cin * KS * G::WROW;      // [C3_CC][ROWS][TWP]
  const int tid = threadIdx.x;
  const int b = blockIdx.z, y0 = blockIdx.y * C3_TH, x0 = blockIdx.x * C3_TW;
  const int HW = H * W;
  for (int e = tid; e < Cin * KS * G::WROW; e += 256) {
    const int ci = e / (KS * G::WROW), r = e - ci * (KS * G::WROW);
    const int ky = r / G::WROW, t = r - ky * G::WROW;
    float v = 0.f;
    if (t < 3 * KS) {
      const int kx = t / 3, c = t - kx * 3;
      const int wy = flip ? KS - 1 - ky : ky, wx = flip ? KS - 1 - kx : kx;
      v = __ldg(w + (size_t)c * w_sc + (size_t)ci * w_sci + wy * KS + wx);
    }
    wS[e] = v;
  }
  const int sx = tid & 15, ry = tid >> 4;             // strip of 4 pixels, row of the tile
  float acc[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
  const float* inb = in + (size_t)b * in_bs;
  for (int c0 = 0; c0 < Cin; c0 += C3_CC) {
    __syncthreads();                                  // previous chunk consumed (and, first time, wS complete)
    for (int e = tid; e < C3_CC * G::ROWS * (C3_TW + KS - 1); e += 256) {
      const int cc = e / (G::ROWS * (C3_TW + KS - 1)), r = e - cc * (G::ROWS * (C3_TW + KS - 1));
      const int yy = r / (C3_TW + KS - 1), xx = r - yy * (C3_TW + KS - 1);
      const int gy = y0 + yy - G::P, gx = x0 + xx - G::P, ci = c0 + cc;
      float v = 0.f;
      if (ci < Cin && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) v = __ldg(inb + (size_t)ci * HW + (size_t)gy * W + gx);
      dS[(cc * G::ROWS + yy) * G::TWP + xx] = v;
    }
    __syncthreads();
    const int nc = min(C3_CC, Cin - c0);
    for (int cc = 0; cc < nc; ++cc) {
      const float* wrow = wS + (size_t)(c0 + cc) * KS * G::WROW;
      const float* drow = dS + (cc * G::ROWS + ry) * G::TWP + 4 * sx;
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
        float d[8];
        const float4 d0 = *reinterpret_cast<const float4*>(drow + ky * G::TWP);
        const float4 d1 = *reinterpret_cast<const float4*>(drow + ky * G::TWP + 4);
        d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
        float wv[G::WROW];
#pragma unroll
        for (int q = 0; q < G::WROW / 4; ++q) {
          const float4 t4 = *reinterpret_cast<const float4*>(wrow + ky * G::WROW + 4 * q);
          wv[4 * q] = t4.x; wv[4 * q + 1] = t4.y; wv[4 * q + 2] = t4.z; wv[4 * q + 3] = t4.w;
        }
#pragma unroll
        for (int kx = 0; kx < KS; ++kx)
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[c][i] = fmaf(wv[kx * 3 + c], d[i + kx], acc[c][i]);
      }
    }
  }
  const int y = y0 + ry, xs = x0 + 4 * sx;
  if (y >= H || xs >= W) return;
  const bool vec = (W % 4 == 0) && (out_bs % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                   (res == nullptr || (res_bs % 4 == 0 && (reinterpret_cast<uintptr_t>(res) & 15) == 0));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const size_t o = (size_t)c * HW + (size_t)y * W + xs;
    float* op = out + (size_t)b * out_bs + o;
    const float* rp = res ? res + (size_t)b * res_bs + o : nullptr;
    if (vec) {                                        // W % 4 == 0 and xs % 4 == 0: the whole strip is inside the row
      float4 v = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
      if (rp) {
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(rp));
        v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
      }
      *reinterpret_cast<float4*>(op) = v;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (xs + i < W) op[i] = acc[c][i] + (rp ? __ldg(rp + i) : 0.f);
    }
  }
}

template <int KS>
static int launch_conv_to3(const float* in, int64_t in_bs, const float* w, int w_sc, int w_sci, int flip, float* out, int64_t out_bs,
                           const float* res, int64_t res_bs, int B, int Cin, int H, int W, cudaStream_t st) {
  using G = C3Geom<KS>;
  const size_t smem = ((size_t)Cin * KS * G::WROW + G::CHUNK) * sizeof(float);
  RCOT_REQUIRE(smem <= 200 * 1024, "conv_to3: %d input channels need %zu bytes of shared memory", Cin, smem);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_to3_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv_to3: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr = smem;
  }
  dim3 grid(cdiv(W, C3_TW), cdiv(H, C3_TH), B);
  conv_to3_kernel<KS><<<grid, 256, smem, st>>>(in, in_bs, w, w_sc, w_sci, flip, out, out_bs, res, res_bs, Cin, H, W);
  return check_launch("conv_to3");
}


// ---------------------------------------------------------------------------------------------------------------
// conv_from3: out[b, o, y, x] = act( sum_{c < 3, ky, kx} in[b, c, y + ky - p, x + kx - p] * W[o, c, ky, kx] + bias[o] )
// (patch_embed 3 -> 48 k3, F_net features.0 3 -> 64 k5 with bias + LeakyReLU, or its bias-free tangent pass with the
// LeakyReLU-derivative mask of a previous forward).  K = 27 / 75 is two or three tcgen05 K chunks of mostly padding; here
// a thread keeps the 3 x k x (4 + k - 1) input window of its 1 x 4 pixel strip in REGISTERS and walks the output
// channels: per channel 3k^2 broadcast weights from shared memory, 12 k^2 FMAs, one 16-byte store.
template <int KS>
__global__ void __launch_bounds__(256) conv_from3_kernel(const float* __restrict__ in, int64_t in_bs, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int64_t out_bs,
                                                         const float* __restrict__ mask_y, int64_t mask_bs, int act, float slope,
                                                         int Cout, int H, int W) {
  using G = C3Geom<KS>;
  constexpr int WIN = 4 + KS - 1;                     // window columns of a strip
  constexpr int WPO = (3 * KS * KS + 3) / 4 * 4;      // weights per output channel, padded to 16 bytes
  extern __shared__ __align__(16) float c3sm[];
  float* wS = c3sm;                                   // [Cout][WPO]
  float* dS = c3sm + (size_t)Cout * WPO;              // [3][ROWS][TWP]
  const int tid = threadIdx.x;
  const int b = blockIdx.z, y0 = blockIdx.y * C3_TH, x0 = blockIdx.x * C3_TW;
  const int HW = H * W;
  for (int e = tid; e < Cout * WPO; e += 256) {
    const int o = e / WPO, t = e - o * WPO;
    wS[e] = t < 3 * KS * KS ? __ldg(w + (size_t)o * 3 * KS * KS + t) : 0.f;
  }
  const float* inb = in + (size_t)b * in_bs;
  for (int e = tid; e < 3 * G::ROWS * (C3_TW + KS - 1); e += 256) {
    const int c = e / (G::ROWS * (C3_TW + KS - 1)), r = e - c * (G::ROWS * (C3_TW + KS - 1));
    const int yy = r / (C3_TW + KS - 1), xx = r - yy * (C3_TW + KS - 1);
    const int gy = y0 + yy - G::P, gx = x0 + xx - G::P;
    float v = 0.f;
    if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) v = __ldg(inb + (size_t)c * HW + (size_t)gy * W + gx);
    dS[(c * G::ROWS + yy) * G::TWP + xx] = v;
  }
  __syncthreads();
  const int sx = tid & 15, ry = tid >> 4;
  float d[3][KS][WIN];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const float* drow = dS + (c * G::ROWS + ry + ky) * G::TWP + 4 * sx;
      const float4 d0 = *reinterpret_cast<const float4*>(drow);
      const float4 d1 = *reinterpret_cast<const float4*>(drow + 4);
      const float t8[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < WIN; ++i) d[c][ky][i] = t8[i];
    }
  const int y = y0 + ry, xs = x0 + 4 * sx;
  if (y >= H || xs >= W) return;
  const bool vec = (W % 4 == 0) && (out_bs % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                   (mask_y == nullptr || (mask_bs % 4 == 0 && (reinterpret_cast<uintptr_t>(mask_y) & 15) == 0));
  const size_t pix = (size_t)y * W + xs;
  for (int o = 0; o < Cout; ++o) {
    const float* wo = wS + (size_t)o * WPO;
    float acc[4];
    const float bv = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = bv;
#pragma unroll
    for (int q = 0; q < WPO / 4; ++q) {
      const float4 w4 = *reinterpret_cast<const float4*>(wo + 4 * q);
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t = 4 * q + k;                      // compile time: (c, ky, kx) of this weight
        if (t < 3 * KS * KS) {
          const int c = t / (KS * KS), ky = (t / KS) % KS, kx = t % KS;
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv[k], d[c][ky][i + kx], acc[i]);
        }
      }
    }
    float* op = out + (size_t)b * out_bs + (size_t)o * HW + pix;
    const float* mp = mask_y ? mask_y + (size_t)b * mask_bs + (size_t)o * HW + pix : nullptr;
    if (vec) {
      float4 v = make_float4(acc[0], acc[1], acc[2], acc[3]);
      if (mp) {
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mp));
        v.x *= m4.x > 0.f ? 1.f : slope; v.y *= m4.y > 0.f ? 1.f : slope;
        v.z *= m4.z > 0.f ? 1.f : slope; v.w *= m4.w > 0.f ? 1.f : slope;
      } else if (act) {
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
      }
      *reinterpret_cast<float4*>(op) = v;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (xs + i < W) {
          float v = acc[i];
          if (mp) v *= __ldg(mp + i) > 0.f ? 1.f : slope;
          else if (act) v = v > 0.f ? v : v * slope;
          op[i] = v;
        }
    }
  }
}

template <int KS>
static int launch_conv_from3(const float* in, int64_t in_bs, const float* w, const float* bias, float* out, int64_t out_bs,
                             const float* mask_y, int64_t mask_bs, int act, float slope, int B, int Cout, int H, int W,
                             cudaStream_t st) {
  using G = C3Geom<KS>;
  constexpr int WPO = (3 * KS * KS + 3) / 4 * 4;
  const size_t smem = ((size_t)Cout * WPO + 3 * G::ROWS * G::TWP) * sizeof(float);
  RCOT_REQUIRE(smem <= 200 * 1024, "conv_from3: %d output channels need %zu bytes of shared memory", Cout, smem);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_from3_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv_from3: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr = smem;
  }
  dim3 grid(cdiv(W, C3_TW), cdiv(H, C3_TH), B);
  conv_from3_kernel<KS><<<grid, 256, smem, st>>>(in, in_bs, w, bias, out, out_bs, mask_y, mask_bs, act, slope, Cout, H, W);
  return check_launch("conv_from3");
}


// ---------------------------------------------------------------------------------------------------------------
// conv3_wgrad: weight gradient of a conv with three channels on one side (stride 1, pad (k-1)/2):
//     acc(cm, c, ty, tx) = sum_{b, y, x} three[b, c, y, x] * many[b, cm, y + ty - p, x + tx - p]
//   conv that ENDS in 3 channels (output conv; weight [3, Cm, k, k]):  many = the conv's input, three = dL/dy,
//       dW[c, cm, ty, tx] += acc(cm, c, ty, tx)
//   conv that STARTS from 3 channels (patch_embed, F_net features.0; weight [Cm, 3, k, k]):  many = dL/dy, three = the input,
//       dW[cm, c, k-1-ty, k-1-tx] += acc(cm, c, ty, tx)
// As a pixel-as-K GEMM the 3-channel operand was padded to a 128-row MMA tile (0.6 ms for the output conv at 128x128,
// batch 32).  Here a CTA owns (32 channels of `many`, one image, a band of 32 rows); thread = (channel, ty) keeps its
// 3k sums in registers and walks the band row by row: the k rows of `many` it needs live in a rolling shared-memory
// buffer (one new row per step, odd row stride = conflict-free across channels), the three-channel row as one float4 per
// pixel (broadcast loads); per pixel 1 + 1 shared loads for 3k FMAs.  One atomicAdd per sum and CTA at the end.
// MEASURED (B200, 128x128, batch 32): 442 us per launch on average -- slower in total than the GEMM launches it replaces
// (2.66 vs 1.81 ms per step): small CTAs walking a serial pixel loop are latency-bound.  Opt-in (RCOT_DIRECT_WGRAD3=1).
constexpr int W3_CH = 32, W3_BAND = 32;

template <int KS>
__global__ void __launch_bounds__(W3_CH* KS) conv3_wgrad_kernel(const float* __restrict__ many, int64_t many_bs,
                                                                const float* __restrict__ three, int64_t three_bs,
                                                                float* __restrict__ dw, int from3, int Cm, int H, int W) {
  constexpr int P = (KS - 1) / 2;
  extern __shared__ __align__(16) float w3sm[];
  const int RS = (W + 2 * P) | 1;                      // odd row stride (floats)
  float* mS = w3sm;                                    // [KS slots][W3_CH][RS]
  float4* tS = reinterpret_cast<float4*>(w3sm + ((size_t)KS * W3_CH * RS + 3) / 4 * 4);   // [W] (t0, t1, t2, 0)
  const int tid = threadIdx.x, nthr = W3_CH * KS;
  const int cm_l = tid % W3_CH, ty = tid / W3_CH;
  const int c0 = blockIdx.x * W3_CH, b = blockIdx.y, y0 = blockIdx.z * W3_BAND;
  const int HW = H * W;
  const float* mb = many + (size_t)b * many_bs;
  const float* tb = three + (size_t)b * three_bs;
  auto load_row = [&](int yy) {                        // row yy of the 32 channels into slot yy mod KS (zeros outside)
    const int slot = ((yy % KS) + KS) % KS;
    float* dst = mS + (size_t)slot * W3_CH * RS;
    const bool in_y = (unsigned)yy < (unsigned)H;
    for (int e = tid; e < W3_CH * (W + 2 * P); e += nthr) {
      const int ch = e / (W + 2 * P), xx = e - ch * (W + 2 * P);
      const int gx = xx - P, cm = c0 + ch;
      float v = 0.f;
      if (in_y && cm < Cm && (unsigned)gx < (unsigned)W) v = __ldg(mb + (size_t)cm * HW + (size_t)yy * W + gx);
      dst[ch * RS + xx] = v;
    }
  };
  float acc[KS][3];
#pragma unroll
  for (int i = 0; i < KS; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
  for (int yy = y0 - P; yy < y0 + P; ++yy) load_row(yy);          // rows y0-P .. y0+P-1
  const int yend = min(y0 + W3_BAND, H);
  for (int y = y0; y < yend; ++y) {
    __syncthreads();                                   // previous row consumed
    load_row(y + P);
    for (int x = tid; x < W; x += nthr)
      tS[x] = make_float4(__ldg(tb + (size_t)y * W + x), __ldg(tb + (size_t)HW + (size_t)y * W + x),
                          __ldg(tb + 2 * (size_t)HW + (size_t)y * W + x), 0.f);
    __syncthreads();
    const int yy = y + ty - P;
    const float* mrow = mS + (size_t)(((yy % KS) + KS) % KS) * W3_CH * RS + cm_l * RS;   // index xx = x + tx
    float win[KS];
#pragma unroll
    for (int i = 0; i < KS - 1; ++i) win[i + 1] = mrow[i];
    for (int x = 0; x < W; ++x) {
#pragma unroll
      for (int i = 0; i < KS - 1; ++i) win[i] = win[i + 1];
      win[KS - 1] = mrow[x + KS - 1];
      const float4 t4 = tS[x];
#pragma unroll
      for (int tx = 0; tx < KS; ++tx) {
        acc[tx][0] = fmaf(t4.x, win[tx], acc[tx][0]);
        acc[tx][1] = fmaf(t4.y, win[tx], acc[tx][1]);
        acc[tx][2] = fmaf(t4.z, win[tx], acc[tx][2]);
      }
    }
  }
  const int cm = c0 + cm_l;
  if (cm >= Cm) return;
#pragma unroll
  for (int tx = 0; tx < KS; ++tx)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = from3 ? (((size_t)cm * 3 + c) * KS + (KS - 1 - ty)) * KS + (KS - 1 - tx)
                             : (((size_t)c * Cm + cm) * KS + ty) * KS + tx;
      atomicAdd(dw + o, acc[tx][c]);
    }
}

template <int KS>
static int launch_conv3_wgrad(const float* many, int64_t many_bs, const float* three, int64_t three_bs, float* dw, int from3, int B,
                              int Cm, int H, int W, cudaStream_t st) {
  constexpr int P = (KS - 1) / 2;
  const int RS = (W + 2 * P) | 1;
  const size_t smem = (((size_t)KS * W3_CH * RS + 3) / 4 * 4) * sizeof(float) + (size_t)W * sizeof(float4);
  RCOT_REQUIRE(smem <= 200 * 1024, "conv3_wgrad: rows of %d pixels need %zu bytes of shared memory", W, smem);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv3_wgrad_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv3_wgrad: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr = smem;
  }
  dim3 grid(cdiv(Cm, W3_CH), B, cdiv(H, W3_BAND));
  conv3_wgrad_kernel<KS><<<grid, W3_CH * KS, smem, st>>>(many, many_bs, three, three_bs, dw, from3, Cm, H, W);
  return check_launch("conv3_wgrad");
}

}  // namespace rcot

using namespace rcot;

extern "C" int rcot_conv_to3(const float* in, int64_t in_bs, const float* weight, int dgrad, float* out, int64_t out_bs,
                             const float* residual, int64_t res_bs, int B, int Cin, int H, int W, int ks, rcot_stream_t st) {
  RCOT_REQUIRE(in && weight && out && B > 0 && B <= 65535 && Cin > 0 && H > 0 && W > 0, "conv_to3: bad arguments");
  RCOT_REQUIRE(ks == 3 || ks == 5, "conv_to3: kernel size 3 or 5 (got %d)", ks);
  RCOT_REQUIRE(cdiv(H, C3_TH) <= 65535, "conv_to3: image too tall");
  // forward: weight [3, Cin, k, k]; data gradient: weight [Cin(= the conv's Cout), 3, k, k], taps flipped
  const int kk = ks * ks;
  const int w_sc = dgrad ? kk : Cin * kk, w_sci = dgrad ? 3 * kk : kk;
  if (ks == 3)
    return launch_conv_to3<3>(in, in_bs, weight, w_sc, w_sci, dgrad ? 1 : 0, out, out_bs, residual, res_bs, B, Cin, H, W, (cudaStream_t)st);
  return launch_conv_to3<5>(in, in_bs, weight, w_sc, w_sci, dgrad ? 1 : 0, out, out_bs, residual, res_bs, B, Cin, H, W, (cudaStream_t)st);
}

extern "C" int rcot_conv_from3(const float* in, int64_t in_bs, const float* weight, const float* bias, float* out, int64_t out_bs,
                               const float* mask_y, int64_t mask_bs, int act, float slope, int B, int Cout, int H, int W, int ks,
                               rcot_stream_t st) {
  RCOT_REQUIRE(in && weight && out && B > 0 && B <= 65535 && Cout > 0 && H > 0 && W > 0, "conv_from3: bad arguments");
  RCOT_REQUIRE(ks == 3 || ks == 5, "conv_from3: kernel size 3 or 5 (got %d)", ks);
  RCOT_REQUIRE(!(mask_y && (bias || act)), "conv_from3: the mask epilogue (tangent pass) takes neither bias nor activation");
  RCOT_REQUIRE(cdiv(H, C3_TH) <= 65535, "conv_from3: image too tall");
  if (ks == 3)
    return launch_conv_from3<3>(in, in_bs, weight, bias, out, out_bs, mask_y, mask_bs, act, slope, B, Cout, H, W, (cudaStream_t)st);
  return launch_conv_from3<5>(in, in_bs, weight, bias, out, out_bs, mask_y, mask_bs, act, slope, B, Cout, H, W, (cudaStream_t)st);
}

extern "C" int rcot_conv3_wgrad(const float* many, int64_t many_bs, const float* three, int64_t three_bs, float* dw, int from3,
                                int B, int Cm, int H, int W, int ks, rcot_stream_t st) {
  RCOT_REQUIRE(many && three && dw && B > 0 && B <= 65535 && Cm > 0 && H > 0 && W > 0, "conv3_wgrad: bad arguments");
  RCOT_REQUIRE(ks == 3 || ks == 5, "conv3_wgrad: kernel size 3 or 5 (got %d)", ks);
  RCOT_REQUIRE(cdiv(H, W3_BAND) <= 65535, "conv3_wgrad: image too tall");
  if (ks == 3) return launch_conv3_wgrad<3>(many, many_bs, three, three_bs, dw, from3 ? 1 : 0, B, Cm, H, W, (cudaStream_t)st);
  return launch_conv3_wgrad<5>(many, many_bs, three, three_bs, dw, from3 ? 1 : 0, B, Cm, H, W, (cudaStream_t)st);
}
