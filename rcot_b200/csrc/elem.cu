// elem.cu -- CUDA-core (HBM-bound) pieces of the Restormer block and its backward:
//   per-pixel LayerNorm statistics and LayerNorm backward   (Net_Restormer.py:173-200)
//   depthwise 3x3 stencil: plain / transposed / GELU-gated / gate backward / weight gradient
//                                                           (Net_Restormer.py:26,75,81-83)
//   pixel (un)shuffle, axpby, per-channel sums (bias gradients)
// Layout everywhere: fp32 NCHW, per-image block contiguous, `*_bs` = batch stride in elements.
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include <stdlib.h>

namespace rcot {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ LayerNorm statistics
// One thread per pixel, loop over channels (coalesced across the warp). Shifted single pass:
// var = E[(x-s)^2] - (E[x-s])^2 with s = x[0], which is as accurate as two passes here.
__global__ void __launch_bounds__(256)
    ln_stats_kernel(const float* __restrict__ x, int64_t x_bs, int C, int HW, float2* __restrict__ stats) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= HW) return;
  const float* xp = x + (size_t)b * x_bs + p;
  const float s = __ldg(xp);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
  for (int c = 0; c < C; ++c) {
    const float d = __ldg(xp + (size_t)c * HW) - s;
    s1 += d;
    s2 = fmaf(d, d, s2);
  }
  const float inv = 1.f / (float)C;
  const float m = s1 * inv;
  const float var = fmaxf(s2 * inv - m * m, 0.f);
  stats[(size_t)b * HW + p] = make_float2(s + m, 1.0f / sqrtf(var + 1e-5f));
}

// Small planes (levels 3-4: 32 images x 256 or 1024 pixels fill only 32-128 CTAs of the one-thread-per-pixel kernel, each
// thread walking up to 384 channels in dependent batches: 25-32 us of pure latency): a CTA takes 32 pixels, its 8 warps
// split the channels and combine their shifted sums through shared memory.
__global__ void __launch_bounds__(256)
    ln_stats_split_kernel(const float* __restrict__ x, int64_t x_bs, int C, int HW, float2* __restrict__ stats) {
  __shared__ float part[2][8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + lane;
  const int b = blockIdx.y;
  const bool valid = p < HW;
  const float* xp = x + (size_t)b * x_bs + (valid ? p : 0);
  const float s = __ldg(xp);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
  for (int c = w; c < C; c += 8) {
    const float d = __ldg(xp + (size_t)c * HW) - s;
    s1 += d;
    s2 = fmaf(d, d, s2);
  }
  part[0][w][lane] = s1;
  part[1][w][lane] = s2;
  __syncthreads();
  if (w == 0 && valid) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      t1 += part[0][k][lane];
      t2 += part[1][k][lane];
    }
    const float inv = 1.f / (float)C;
    const float m = t1 * inv;
    const float var = fmaxf(t2 * inv - m * m, 0.f);
    stats[(size_t)b * HW + p] = make_float2(s + m, 1.0f / sqrtf(var + 1e-5f));
  }
}

// ------------------------------------------------------------------ LayerNorm forward (stand-alone module)
// y = (x - mu) * rstd * gamma + beta per pixel over C (Net_Restormer.py:186-189), statistics written for the backward.
// Inside the blocks LayerNorm is a prologue of the consuming GEMM; this kernel backs `LayerNorm.forward` on its own.
__global__ void __launch_bounds__(256)
    ln_fwd_kernel(const float* __restrict__ x, int64_t x_bs, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ y, int64_t y_bs, int C, int HW,
                  float2* __restrict__ stats) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= HW) return;
  const float* xp = x + (size_t)b * x_bs + p;
  const float s = __ldg(xp);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
  for (int c = 0; c < C; ++c) {
    const float d = __ldg(xp + (size_t)c * HW) - s;
    s1 += d;
    s2 = fmaf(d, d, s2);
  }
  const float inv = 1.f / (float)C;
  const float m = s1 * inv;
  const float var = fmaxf(s2 * inv - m * m, 0.f);
  const float mu = s + m, rstd = 1.0f / sqrtf(var + 1e-5f);
  stats[(size_t)b * HW + p] = make_float2(mu, rstd);
  float* yp = y + (size_t)b * y_bs + p;
#pragma unroll 8
  for (int c = 0; c < C; ++c)
    yp[(size_t)c * HW] = (__ldg(xp + (size_t)c * HW) - mu) * rstd * __ldg(gamma + c) + __ldg(beta + c);
}

// ------------------------------------------------------------------ LayerNorm backward
// g = dz*gamma ; dx = [dy +] rstd*(g - mean_c(g) - xhat*mean_c(g*xhat)) ; dgamma += sum dz*xhat ;
// dbeta += sum dz.  CTA = 32 pixels (lanes) x 8 channel groups (warps); warp w owns channels w, w+8, ...
// so the per-channel sums over pixels are warp shuffles into shared slots no other warp touches; the
// per-pixel sums over channels are combined across the 8 warps through shared memory.  A CTA walks
// `groups` consecutive 32-pixel groups before it flushes its channel sums with one atomicAdd each.
template <int NCH, int MINB = 2, bool EARLYDY = false, int NW = 8>   // NCH > 0: C == NW*NCH and the (dz, xhat) values of the first sweep stay in registers
__global__ void __launch_bounds__(32 * NW, MINB)
    ln_bwd_kernel(const float* __restrict__ dz, int64_t dz_bs, const float* __restrict__ x, int64_t x_bs,
                  const float2* __restrict__ stats, const float* __restrict__ gamma, const float* dy, int64_t dy_bs,
                  float* dx, int64_t dx_bs, float* __restrict__ dgamma, float* __restrict__ dbeta, int C, int HW,
                  int groups) {
  extern __shared__ float sacc[];  // [2*C] channel sums, then [2][NW][32] pixel partials
  float* spix = sacc + 2 * C;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  constexpr int R = NCH > 0 ? NCH : 1;
  float ag[R], ab[R];   // per-thread partial channel sums over this CTA's pixel groups (register path)
#pragma unroll
  for (int k = 0; k < R; ++k) {
    ag[k] = 0.f;
    ab[k] = 0.f;
  }
  for (int gi = 0; gi < groups; ++gi) {
    const int p = (blockIdx.x * groups + gi) * 32 + lane;
    const bool valid = p < HW;
    const float* xp = x + (size_t)b * x_bs + p;
    const float* dzp = dz + (size_t)b * dz_bs + p;
    float mu = 0.f, rstd = 0.f;
    if (valid) {
      const float2 st = stats[(size_t)b * HW + p];
      mu = st.x;
      rstd = st.y;
    }
    float sg = 0.f, sgx = 0.f;
    float rd[R], rx[R];
    if (NCH > 0) {
#pragma unroll
      for (int k = 0; k < R; ++k) {   // all loads first: 2*NCH independent requests per thread
        const int c = wy + NW * k;
        rd[k] = valid ? __ldg(dzp + (size_t)c * HW) : 0.f;
        rx[k] = valid ? __ldg(xp + (size_t)c * HW) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int c = wy + NW * k;
        const float xh = valid ? (rx[k] - mu) * rstd : 0.f;
        rx[k] = xh;
        const float g = rd[k] * __ldg(gamma + c);
        sg += g;
        sgx = fmaf(g, xh, sgx);
        ag[k] = fmaf(rd[k], xh, ag[k]);   // reduced over the 32 pixels of the warp once, after the group loop
        ab[k] += rd[k];
      }
    } else {
      for (int c = wy; c < C; c += NW) {
        float d = 0.f, xh = 0.f;
        if (valid) {
          d = __ldg(dzp + (size_t)c * HW);
          xh = (__ldg(xp + (size_t)c * HW) - mu) * rstd;
        }
        const float g = d * __ldg(gamma + c);
        sg += g;
        sgx = fmaf(g, xh, sgx);
        const float wg = warp_sum(d * xh);
        const float wb = warp_sum(d);
        if (lane == 0) {
          sacc[c] += wg;
          sacc[C + c] += wb;
        }
      }
    }
    float ry[R];
    if (NCH > 0 && EARLYDY) {          // the residual rows travel while the CTA meets at the barrier below
      const float* dyp = (dy && valid) ? dy + (size_t)b * dy_bs + p : nullptr;
#pragma unroll
      for (int k = 0; k < R; ++k) ry[k] = dyp ? dyp[(size_t)(wy + NW * k) * HW] : 0.f;
    }
    spix[wy * 32 + lane] = sg;
    spix[NW * 32 + wy * 32 + lane] = sgx;
    __syncthreads();
    float mg = 0.f, mgx = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      mg += spix[w * 32 + lane];
      mgx += spix[NW * 32 + w * 32 + lane];
    }
    const float inv = 1.f / (float)C;
    mg *= inv;
    mgx *= inv;
    if (valid) {
      const float* dyp = dy ? dy + (size_t)b * dy_bs + p : nullptr;
      float* dxp = dx + (size_t)b * dx_bs + p;
      if (NCH > 0) {
        if (!EARLYDY) {
#pragma unroll
          for (int k = 0; k < R; ++k) ry[k] = dyp ? dyp[(size_t)(wy + NW * k) * HW] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int c = wy + NW * k;
          dxp[(size_t)c * HW] = rstd * (rd[k] * __ldg(gamma + c) - mg - rx[k] * mgx) + ry[k];
        }
      } else {
        for (int c = wy; c < C; c += NW) {
          const float d = __ldg(dzp + (size_t)c * HW);
          const float xh = (__ldg(xp + (size_t)c * HW) - mu) * rstd;
          float r = rstd * (d * __ldg(gamma + c) - mg - xh * mgx);
          if (dyp) r += dyp[(size_t)c * HW];
          dxp[(size_t)c * HW] = r;
        }
      }
    }
    __syncthreads();
  }
  if (NCH > 0) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float wg = warp_sum(ag[k]);
      const float wb = warp_sum(ab[k]);
      if (lane == 0) {          // channel wy + NW k belongs to this warp alone
        sacc[wy + NW * k] = wg;
        sacc[C + wy + NW * k] = wb;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, sacc[i]);
    atomicAdd(dbeta + i, sacc[C + i]);
  }
}

// ------------------------------------------------------------------ depthwise 3x3
__device__ __forceinline__ float gelu_erf(float a) { return 0.5f * a * (1.f + erff(a * 0.70710678118654752f)); }
// d/da gelu(a) = Phi(a) + a*phi(a)
__device__ __forceinline__ float gelu_erf_grad(float a) {
  return 0.5f * (1.f + erff(a * 0.70710678118654752f)) + a * 0.39894228040143268f * __expf(-0.5f * a * a);
}

// Each thread produces 4 horizontally adjacent outputs of one (image, channel) plane: three rows of
// 4+2 inputs are fetched once (one aligned float4 + two edge scalars per row) and reused by the 9 taps.
// W % 4 == 0 is required by the launcher (all feature maps here have W in {4,8,...,256}).
struct Row6 {
  float v[6];  // x0-1 .. x0+4
};
__device__ __forceinline__ Row6 load_row6(const float* __restrict__ plane, int y, int x0, int H, int W) {
  Row6 r;
  if ((unsigned)y >= (unsigned)H) {
#pragma unroll
    for (int i = 0; i < 6; ++i) r.v[i] = 0.f;
    return r;
  }
  const float* p = plane + (size_t)y * W + x0;
  const float4 m = __ldg(reinterpret_cast<const float4*>(p));
  r.v[0] = x0 > 0 ? __ldg(p - 1) : 0.f;
  r.v[1] = m.x;
  r.v[2] = m.y;
  r.v[3] = m.z;
  r.v[4] = m.w;
  r.v[5] = x0 + 4 < W ? __ldg(p + 4) : 0.f;
  return r;
}
// out[j] = sum_{ky,kx} w[ky*3+kx] * in(y+ky-1, x0+j+kx-1), j = 0..3
__device__ __forceinline__ void stencil4(const float* __restrict__ plane, const float* w, int y, int x0, int H, int W,
                                         float* out) {
#pragma unroll
  for (int j = 0; j < 4; ++j) out[j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const Row6 r = load_row6(plane, y + ky - 1, x0, H, W);
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = fmaf(r.v[j + kx], w[ky * 3 + kx], out[j]);
  }
}

// Two vertically adjacent output rows (y, y+1) from the four input rows y-1 .. y+2: 8 outputs per
// 4 row fetches instead of 6.
__device__ __forceinline__ void stencil4x2(const float* __restrict__ plane, const float* w, int y, int x0, int H, int W,
                                           float* o0, float* o1) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o0[j] = 0.f;
    o1[j] = 0.f;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const Row6 v = load_row6(plane, y + r - 1, x0, H, W);
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (r < 3) o0[j] = fmaf(v.v[j + kx], w[r * 3 + kx], o0[j]);
        if (r > 0) o1[j] = fmaf(v.v[j + kx], w[(r - 1) * 3 + kx], o1[j]);
      }
  }
}

// mode 0: out[ch] = dw(in[ch])                (flip=1: transposed = data gradient)
//         optional sumsq[b*nsq + ch] += sum_p out^2 for ch < nsq   (MDTA row norms of q and k)
// mode 1: out[j] = gelu(dw(in[j])) * dw(in[j+hid]),  j < hid            (GDFN gate)
// mode 2: a = dw(in[j]), b = dw(in[j+hid]); out[j] = dg*b*gelu'(a); out[j+hid] = dg*gelu(a);
//         optional g_out[j] = gelu(a)*b                                 (GDFN gate backward)
// Work is flattened over (image, plane, row pair, quad column) so small feature maps still fill the machine;
// each thread produces a 4-wide x 2-high patch (H is even at every level).
__global__ void __launch_bounds__(256) dwconv_kernel(const rcot_dw_params p, const int planes, const int qpp,
                                                     const long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < total;
  const long pl = active ? idx / qpp : 0;
  const int quad = active ? (int)(idx - pl * qpp) : 0;
  const int b = (int)(pl / planes), ch = (int)(pl - (long)b * planes);
  const int HW = p.H * p.W, qw = p.W >> 2;
  const int yp = quad / qw;
  const int y = yp * 2, x0 = (quad - yp * qw) * 4;
  const int pix = y * p.W + x0;
  const float* inb = p.in + (size_t)b * p.in_bs;
  float w0[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w0[i] = __ldg(p.w + ch * 9 + (p.flip ? 8 - i : i));
  if (p.mode == 0) {
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
      stencil4x2(inb + (size_t)ch * HW, w0, y, x0, p.H, p.W, o0, o1);
      float* op = p.out + (size_t)b * p.out_bs + (size_t)ch * HW + pix;
      *reinterpret_cast<float4*>(op) = make_float4(o0[0], o0[1], o0[2], o0[3]);
      *reinterpret_cast<float4*>(op + p.W) = make_float4(o1[0], o1[1], o1[2], o1[3]);
    }
    if (p.sumsq) {  // uniform branch
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) sq += o0[j] * o0[j] + o1[j] * o1[j];
      const bool mine = active && ch < p.nsq;
      if (qpp % 32 == 0) {  // a warp never straddles two planes
        sq = warp_sum(mine ? sq : 0.f);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(p.sumsq + (size_t)b * p.nsq + ch, sq);
      } else if (mine) {
        atomicAdd(p.sumsq + (size_t)b * p.nsq + ch, sq);
      }
    }
    return;
  }
  if (!active) return;
  float w1[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w1[i] = __ldg(p.w + (ch + p.hid) * 9 + i);
  float a[2][4], g[2][4];
  stencil4x2(inb + (size_t)ch * HW, w0, y, x0, p.H, p.W, a[0], a[1]);
  stencil4x2(inb + (size_t)(ch + p.hid) * HW, w1, y, x0, p.H, p.W, g[0], g[1]);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int pr = pix + r * p.W;
    if (p.mode == 1) {
      *reinterpret_cast<float4*>(p.out + (size_t)b * p.out_bs + (size_t)ch * HW + pr) =
          make_float4(gelu_erf(a[r][0]) * g[r][0], gelu_erf(a[r][1]) * g[r][1], gelu_erf(a[r][2]) * g[r][2],
                      gelu_erf(a[r][3]) * g[r][3]);
    } else {
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(p.dg + (size_t)b * p.dg_bs + (size_t)ch * HW + pr));
      const float d[4] = {d4.x, d4.y, d4.z, d4.w};
      float da[4], db[4], gg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ga = gelu_erf(a[r][j]);
        da[j] = d[j] * g[r][j] * gelu_erf_grad(a[r][j]);
        db[j] = d[j] * ga;
        gg[j] = ga * g[r][j];
      }
      float* ob = p.out + (size_t)b * p.out_bs;
      *reinterpret_cast<float4*>(ob + (size_t)ch * HW + pr) = make_float4(da[0], da[1], da[2], da[3]);
      *reinterpret_cast<float4*>(ob + (size_t)(ch + p.hid) * HW + pr) = make_float4(db[0], db[1], db[2], db[3]);
      if (p.g_out)
        *reinterpret_cast<float4*>(p.g_out + (size_t)b * p.g_bs + (size_t)ch * HW + pr) =
            make_float4(gg[0], gg[1], gg[2], gg[3]);
    }
  }
}

// Fallback for feature maps whose width is not a multiple of 4 or whose height is odd (whole-image
// inference at arbitrary sizes, reference tester.py:107): one output per thread, modes 0 and 1 only.
__global__ void __launch_bounds__(256) dwconv_scalar_kernel(const rcot_dw_params p, const int planes, const long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int HW = p.H * p.W;
  const long pl = idx / HW;
  const int pix = (int)(idx - pl * HW);
  const int b = (int)(pl / planes), ch = (int)(pl - (long)b * planes);
  const int y = pix / p.W, x = pix - y * p.W;
  const float* inb = p.in + (size_t)b * p.in_bs;
  auto tap9 = [&](int c, bool flip) {
    const float* plane = inb + (size_t)c * HW;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if ((unsigned)yy >= (unsigned)p.H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if ((unsigned)xx >= (unsigned)p.W) continue;
        const int k = ky * 3 + kx;
        acc = fmaf(__ldg(plane + yy * p.W + xx), __ldg(p.w + c * 9 + (flip ? 8 - k : k)), acc);
      }
    }
    return acc;
  };
  float o = tap9(ch, p.flip != 0);
  if (p.mode == 1) o = gelu_erf(o) * tap9(ch + p.hid, false);
  p.out[(size_t)b * p.out_bs + (size_t)ch * HW + pix] = o;
  if (p.mode == 0 && p.sumsq && ch < p.nsq) atomicAdd(p.sumsq + (size_t)b * p.nsq + ch, o * o);
}

// dW[ch, k] += sum_{b,p} dout[b,ch,p] * in[b,ch,p+off_k].  grid = (chunks, Cn); each CTA strides over
// (image, quad) pairs of its channel, keeps 9 partial sums per thread, reduces and issues 9 atomics.
__global__ void __launch_bounds__(256)
    dw_wgrad_kernel(const float* __restrict__ in, int64_t in_bs, const float* __restrict__ dout, int64_t dout_bs,
                    float* __restrict__ dw, int B, int H, int W) {
  const int HW = H * W, ch = blockIdx.y, qpp = HW / 4;
  const long total = (long)B * qpp;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int b = (int)(e / qpp), pix = (int)(e - (long)b * qpp) * 4;
    const int y = pix / W, x0 = pix - y * W;
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(dout + (size_t)b * dout_bs + (size_t)ch * HW + pix));
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
    const float* plane = in + (size_t)b * in_bs + (size_t)ch * HW;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const Row6 r = load_row6(plane, y + ky - 1, x0, H, W);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[ky * 3 + kx] = fmaf(d[j], r.v[j + kx], acc[ky * 3 + kx]);
    }
  }
  __shared__ float red[9][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[i][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    atomicAdd(dw + ch * 9 + threadIdx.x, s);
  }
}

// Backward of the depthwise conv in ONE pass over its two inputs:
//   din = dw^T(dout)  (same stencil with flipped taps)   and   dW[ch,k] += sum dout * in(shifted by k).
// grid = (chunks, Cn): a CTA owns one channel and strides over (image, quad) pairs.
__global__ void __launch_bounds__(256)
    dw_bwd_kernel(const float* __restrict__ in, int64_t in_bs, const float* __restrict__ dout, int64_t dout_bs,
                  const float* __restrict__ w, float* __restrict__ din, int64_t din_bs, float* __restrict__ dw, int B,
                  int H, int W) {
  const int HW = H * W, ch = blockIdx.y, qw = W >> 2, qpp = (H >> 1) * qw;   // 4 x 2 patches per plane
  const long total = (long)B * qpp;
  float wf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wf[i] = __ldg(w + ch * 9 + 8 - i);   // flipped taps
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int b = (int)(e / qpp), quad = (int)(e - (long)b * qpp);
    const int yp = quad / qw;
    const int y = yp * 2, x0 = (quad - yp * qw) * 4;
    const float* dplane = dout + (size_t)b * dout_bs + (size_t)ch * HW;
    const float* iplane = in + (size_t)b * in_bs + (size_t)ch * HW;
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    float d0[4], d1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {   // dout rows y-1 .. y+2 feed din rows y and y+1
      const Row6 v = load_row6(dplane, y + r - 1, x0, H, W);
      if (r == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) d0[j] = v.v[j + 1];
      }
      if (r == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) d1[j] = v.v[j + 1];
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (r < 3) o0[j] = fmaf(v.v[j + kx], wf[r * 3 + kx], o0[j]);
          if (r > 0) o1[j] = fmaf(v.v[j + kx], wf[(r - 1) * 3 + kx], o1[j]);
        }
    }
    float* dp = din + (size_t)b * din_bs + (size_t)ch * HW + y * W + x0;
    *reinterpret_cast<float4*>(dp) = make_float4(o0[0], o0[1], o0[2], o0[3]);
    *reinterpret_cast<float4*>(dp + W) = make_float4(o1[0], o1[1], o1[2], o1[3]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {   // in rows y-1 .. y+2 against the two centre rows of dout
      const Row6 v = load_row6(iplane, y + r - 1, x0, H, W);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (r < 3) acc[r * 3 + kx] = fmaf(d0[j], v.v[j + kx], acc[r * 3 + kx]);
          if (r > 0) acc[(r - 1) * 3 + kx] = fmaf(d1[j], v.v[j + kx], acc[(r - 1) * 3 + kx]);
        }
    }
  }
  __shared__ float red[9][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[i][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
    atomicAdd(dw + ch * 9 + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ pixel (un)shuffle, r = 2
// inverse=0 (PixelShuffle):   out[b, c, 2y+i, 2x+j] = in[b, 4c+2i+j, y, x]     in: [4C,H,W]  out: [C,2H,2W]
// inverse=1 (PixelUnshuffle): out[b, 4c+2i+j, y, x] = in[b, c, 2y+i, 2x+j]     in: [C,2H,2W] out: [4C,H,W]
// C,H,W always describe the (4C,H,W) side. One thread per element of the (C,2H,2W) side.
__global__ void __launch_bounds__(256)
    pixel_shuffle_kernel(const float* __restrict__ in, int64_t in_bs, float* __restrict__ out, int64_t out_bs, int C,
                         int H, int W, int inverse) {
  const int W2 = 2 * W, H2 = 2 * H;
  const long n = (long)C * H2 * W2;
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (e >= n) return;
  const int c = (int)(e / ((long)H2 * W2));
  const int r = (int)(e - (long)c * H2 * W2);
  const int yy = r / W2, xx = r - yy * W2;
  const size_t small = ((size_t)(4 * c + 2 * (yy & 1) + (xx & 1)) * H + (yy >> 1)) * W + (xx >> 1);
  if (inverse) out[(size_t)b * out_bs + small] = __ldg(in + (size_t)b * in_bs + e);
  else out[(size_t)b * out_bs + e] = __ldg(in + (size_t)b * in_bs + small);
}

// ------------------------------------------------------------------ out = a*x + b*y (per-sample a optional)
__global__ void __launch_bounds__(256)
    axpby_kernel(float* out, int64_t out_bs, const float* x, int64_t x_bs, const float* y, int64_t y_bs, float a,
                 float bcoef, const float* __restrict__ a_vec, int mode, long n) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (e >= n) return;
  float av = a, bv = bcoef;
  if (mode == 1) {  // interpolation: a_vec[b]*x + (1-a_vec[b])*y
    av = __ldg(a_vec + b);
    bv = 1.f - av;
  }
  const float xv = x[(size_t)b * x_bs + e];
  const float yv = y ? y[(size_t)b * y_bs + e] : 0.f;
  out[(size_t)b * out_bs + e] = av * xv + bv * yv;
}

// out[c] += sum_{b,p} x[b,c,p]   (bias gradients).  grid = (chunks, C)
__global__ void __launch_bounds__(256)
    channel_sum_kernel(const float* __restrict__ x, int64_t x_bs, float* __restrict__ out, int B, int HW) {
  const int c = blockIdx.y;
  const long total = (long)B * HW;
  float s = 0.f;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int b = (int)(e / HW), pix = (int)(e - (long)b * HW);
    s += __ldg(x + (size_t)b * x_bs + (size_t)c * HW + pix);
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out + c, t);
  }
}

}  // namespace rcot

namespace rcot {   // dwconv.cu: lean kernels for aligned feature maps; return 0 when a call does not qualify
int dwconv_fast(const rcot_dw_params& p, int planes, cudaStream_t st);
int gdfn_mid_bwd_fast(const float* u, int64_t u_bs, const float* dg, int64_t dg_bs, const float* w, float* du,
                      int64_t du_bs, float* dw, float* g_out, int64_t g_bs, int B, int hid, int H, int W,
                      cudaStream_t st);
int dwconv_bwd_fast(const void* in, int64_t in_bs, const void* dout, int64_t dout_bs, const float* w, void* din,
                    int64_t din_bs, float* dw, int B, int Cn, int H, int W, int bf16, cudaStream_t st);
}  // namespace rcot

using namespace rcot;

extern "C" int rcot_ln_stats(const float* x, int64_t x_bs, int B, int C, int HW, float* stats, rcot_stream_t st) {
  RCOT_REQUIRE(x && stats && B > 0 && C > 0 && HW > 0, "ln_stats: bad arguments");
  RCOT_REQUIRE(B <= 65535, "ln_stats: batch too large");
  dim3 grid(cdiv(HW, 256), B);
  if ((long)grid.x * B < 2 * 148 && C >= 64) {
    dim3 grid2(cdiv(HW, 32), B);
    ln_stats_split_kernel<<<grid2, 256, 0, (cudaStream_t)st>>>(x, x_bs, C, HW, reinterpret_cast<float2*>(stats));
    return check_launch("ln_stats");
  }
  ln_stats_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(x, x_bs, C, HW, reinterpret_cast<float2*>(stats));
  return check_launch("ln_stats");
}

extern "C" int rcot_ln_fwd(const float* x, int64_t x_bs, const float* gamma, const float* beta, float* y, int64_t y_bs,
                           int B, int C, int HW, float* stats, rcot_stream_t st) {
  RCOT_REQUIRE(x && gamma && beta && y && stats && B > 0 && C > 0 && HW > 0, "ln_fwd: bad arguments");
  RCOT_REQUIRE(B <= 65535, "ln_fwd: batch too large");
  dim3 grid(cdiv(HW, 256), B);
  ln_fwd_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(x, x_bs, gamma, beta, y, y_bs, C, HW, reinterpret_cast<float2*>(stats));
  return check_launch("ln_fwd");
}

extern "C" int rcot_ln_bwd(const float* dz, int64_t dz_bs, const float* x, int64_t x_bs, const float* stats,
                           const float* gamma, const float* dy, int64_t dy_bs, float* dx, int64_t dx_bs,
                           float* dgamma, float* dbeta, int B, int C, int HW, rcot_stream_t st) {
  RCOT_REQUIRE(dz && x && stats && gamma && dx && dgamma && dbeta, "ln_bwd: null pointer");
  RCOT_REQUIRE(B > 0 && B <= 65535 && C > 0 && HW > 0, "ln_bwd: bad sizes");
  // enough CTAs for ~8 per SM, at most 8 pixel groups (256 pixels) per CTA
  long pg = cdiv(HW, 32);
  int groups = (int)((pg * B) / (148 * 8));
  if (groups < 1) groups = 1;
  if (groups > 8) groups = 8;
  dim3 grid(cdiv(pg, groups), B);
  const float2* st2 = reinterpret_cast<const float2*>(stats);
  // Measured (scripts/bench_ln.py, B=32): the kernel is latency-bound (ncu at C=96, 8 warps x 12 channels per thread:
  // 128 registers, 24 % warps active, 55 % long-scoreboard stalls).  What paid, per width:
  //   C=48 : 3 CTAs per SM (80 registers) + residual rows requested before the CTA barrier      136 -> 82 us (4.9 TB/s)
  //   C=192: 32 warps x 6 channels per thread (64 registers, no spills; 8 x 24 spilled)          48 -> 44 us
  //   C=384: 32 warps x 12 channels in registers instead of the generic re-reading loop           50 -> 32 us
  //   C=96 : unchanged (8 warps x 12); 16 warps x 6 measured 213 vs 186 us, 3 CTAs per SM spill: 193 us
  // A/B switch RCOT_LN_BWD_VAR=1: the round-1 shapes (8 warps, NCH = C/8) everywhere.
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("RCOT_LN_BWD_VAR");
    forced = e ? atoi(e) : 0;
  }
#define LNBV(NCH, MB, ED, NW_)                                                                                   \
  ln_bwd_kernel<NCH, MB, ED, NW_><<<grid, 32 * NW_, (2 * C + 64 * NW_) * sizeof(float), (cudaStream_t)st>>>(     \
      dz, dz_bs, x, x_bs, st2, gamma, dy, dy_bs, dx, dx_bs, dgamma, dbeta, C, HW, groups)
  if (forced == 1) {
    if (C == 48) LNBV(6, 2, false, 8);
    else if (C == 96) LNBV(12, 2, false, 8);
    else if (C == 192) LNBV(24, 2, false, 8);
    else LNBV(0, 2, false, 8);
  } else {
    if (C == 48) LNBV(6, 3, true, 8);
    else if (C == 96) LNBV(12, 2, false, 8);
    else if (C == 192) LNBV(6, 1, true, 32);
    else if (C == 384) LNBV(12, 1, true, 32);
    else LNBV(0, 2, false, 8);
  }
#undef LNBV
  return check_launch("ln_bwd");
}

extern "C" int rcot_dwconv3x3(const rcot_dw_params* pp, rcot_stream_t st) {
  RCOT_REQUIRE(pp != nullptr, "dwconv3x3: null params");
  const rcot_dw_params& p = *pp;
  RCOT_REQUIRE(p.in && p.w && p.out, "dwconv3x3: null pointer");
  RCOT_REQUIRE(p.B > 0 && p.B <= 65535 && p.Cn > 0 && p.H > 0 && p.W > 0, "dwconv3x3: bad sizes");
  RCOT_REQUIRE(p.mode >= 0 && p.mode <= 2, "dwconv3x3: bad mode %d", p.mode);
  int planes = p.Cn;
  if (p.mode != 0) {
    RCOT_REQUIRE(p.hid > 0 && 2 * p.hid == p.Cn, "dwconv3x3: gate modes need Cn == 2*hid");
    RCOT_REQUIRE(p.flip == 0, "dwconv3x3: gate modes cannot flip");
    if (p.mode == 2) RCOT_REQUIRE(p.dg != nullptr, "dwconv3x3: gate backward needs dg");
    planes = p.hid;
  }
  if (dwconv_fast(p, planes, (cudaStream_t)st)) return check_launch("dwconv3x3");
  RCOT_REQUIRE(!p.bf16, "dwconv3x3: bf16 storage needs the aligned fast path (W %% 4 == 0, H even, 16-byte aligned; got %dx%d)",
               p.H, p.W);
  if (p.W % 4 != 0 || p.H % 2 != 0) {
    RCOT_REQUIRE(p.mode != 2, "dwconv3x3: gate backward needs width %% 4 == 0 and even height, got %dx%d", p.H, p.W);
    const long tot = (long)p.B * planes * p.H * p.W;
    dwconv_scalar_kernel<<<cdiv(tot, 256), 256, 0, (cudaStream_t)st>>>(p, planes, tot);
    return check_launch("dwconv3x3");
  }
  RCOT_REQUIRE(p.in_bs % 4 == 0 && p.out_bs % 4 == 0 && ((uintptr_t)p.in % 16 == 0) && ((uintptr_t)p.out % 16 == 0),
               "dwconv3x3: tensors must be 16-byte aligned");
  const int qpp = p.H * p.W / 8;   // 4 x 2 output patches per plane
  const long total = (long)p.B * planes * qpp;
  dwconv_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)st>>>(p, planes, qpp, total);
  return check_launch("dwconv3x3");
}

extern "C" int rcot_dwconv3x3_wgrad(const float* in, int64_t in_bs, const float* dout, int64_t dout_bs, float* dw,
                                    int B, int Cn, int H, int W, rcot_stream_t st) {
  RCOT_REQUIRE(in && dout && dw && B > 0 && Cn > 0 && Cn <= 65535 && H > 0 && W > 0, "dwconv3x3_wgrad: bad arguments");
  RCOT_REQUIRE(W % 4 == 0 && in_bs % 4 == 0 && dout_bs % 4 == 0, "dwconv3x3_wgrad: width/strides must be multiples of 4");
  long total = (long)B * H * W / 4;
  int chunks = (int)((total + 256 * 8 - 1) / (256 * 8));
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  dim3 grid(chunks, Cn);
  dw_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(in, in_bs, dout, dout_bs, dw, B, H, W);
  return check_launch("dwconv3x3_wgrad");
}

extern "C" int rcot_dwconv3x3_bwd(const float* in, int64_t in_bs, const float* dout, int64_t dout_bs, const float* w,
                                  float* din, int64_t din_bs, float* dw, int B, int Cn, int H, int W,
                                  rcot_stream_t st) {
  RCOT_REQUIRE(in && dout && w && din && dw && B > 0 && Cn > 0 && Cn <= 65535 && H > 0 && W > 0,
               "dwconv3x3_bwd: bad arguments");
  RCOT_REQUIRE(W % 4 == 0 && H % 2 == 0 && in_bs % 4 == 0 && dout_bs % 4 == 0 && din_bs % 4 == 0,
               "dwconv3x3_bwd: width/strides must be multiples of 4 and height even");
  if (dwconv_bwd_fast(in, in_bs, dout, dout_bs, w, din, din_bs, dw, B, Cn, H, W, 0, (cudaStream_t)st))
    return check_launch("dwconv3x3_bwd");
  long total = (long)B * H * W / 8;
  // enough CTAs per channel to fill the machine, at least ~4 quads per thread
  long want = (148L * 8 + Cn - 1) / Cn;
  long maxc = (total + 256 * 2 - 1) / (256 * 2);
  int chunks = (int)(want < maxc ? want : maxc);
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, Cn);
  dw_bwd_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(in, in_bs, dout, dout_bs, w, din, din_bs, dw, B, H, W);
  return check_launch("dwconv3x3_bwd");
}

extern "C" int rcot_dwconv3x3_bwd_t(const void* in, int64_t in_bs, const void* dout, int64_t dout_bs, const float* w,
                                    void* din, int64_t din_bs, float* dw, int B, int Cn, int H, int W, int bf16,
                                    rcot_stream_t st) {
  if (!bf16)
    return rcot_dwconv3x3_bwd(reinterpret_cast<const float*>(in), in_bs, reinterpret_cast<const float*>(dout), dout_bs, w,
                              reinterpret_cast<float*>(din), din_bs, dw, B, Cn, H, W, st);
  RCOT_REQUIRE(in && dout && w && din && dw && B > 0 && Cn > 0 && Cn <= 65535 && H > 0 && W > 0,
               "dwconv3x3_bwd: bad arguments");
  RCOT_REQUIRE(W % 4 == 0 && H % 2 == 0 && in_bs % 4 == 0 && dout_bs % 4 == 0 && din_bs % 4 == 0,
               "dwconv3x3_bwd: width/strides must be multiples of 4 and height even");
  RCOT_REQUIRE(dwconv_bwd_fast(in, in_bs, dout, dout_bs, w, din, din_bs, dw, B, Cn, H, W, 1, (cudaStream_t)st) == 1,
               "dwconv3x3_bwd: bf16 storage needs the aligned fast path");
  return check_launch("dwconv3x3_bwd");
}

extern "C" int rcot_gdfn_mid_bwd(const float* u, int64_t u_bs, const float* dg, int64_t dg_bs, const float* w, float* du,
                                 int64_t du_bs, float* dw, float* g_out, int64_t g_bs, int B, int hid, int H, int W,
                                 rcot_stream_t st) {
  RCOT_REQUIRE(u && dg && w && du && dw && B > 0 && hid > 0 && H > 0 && W > 0, "gdfn_mid_bwd: bad arguments");
  RCOT_REQUIRE(W % 32 == 0 && H % 4 == 0 && B <= 65535 && hid <= 65535,
               "gdfn_mid_bwd: needs width %% 32 == 0 and height %% 4 == 0 (got %dx%d)", H, W);
  RCOT_REQUIRE(u_bs % 4 == 0 && du_bs % 4 == 0 && dg_bs % 4 == 0 && ((uintptr_t)u % 16 == 0) &&
                   ((uintptr_t)du % 16 == 0) && ((uintptr_t)dg % 16 == 0) &&
                   (!g_out || (g_bs % 4 == 0 && (uintptr_t)g_out % 16 == 0)),
               "gdfn_mid_bwd: u, dg, du and g_out must be 16-byte aligned");
  RCOT_REQUIRE(gdfn_mid_bwd_fast(u, u_bs, dg, dg_bs, w, du, du_bs, dw, g_out, g_bs, B, hid, H, W, (cudaStream_t)st) == 1,
               "gdfn_mid_bwd: geometry not supported");
  return check_launch("gdfn_mid_bwd");
}

extern "C" int rcot_pixel_shuffle(const float* in, int64_t in_bs, float* out, int64_t out_bs, int B, int C, int H,
                                  int W, int inverse, rcot_stream_t st) {
  RCOT_REQUIRE(in && out && B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0, "pixel_shuffle: bad arguments");
  const long n = (long)C * 4 * H * W;
  dim3 grid(cdiv(n, 256), B);
  pixel_shuffle_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(in, in_bs, out, out_bs, C, H, W, inverse);
  return check_launch("pixel_shuffle");
}

extern "C" int rcot_axpby(float* out, int64_t out_bs, const float* x, int64_t x_bs, const float* y, int64_t y_bs,
                          float a, float b, const float* a_vec, int B, int64_t n, rcot_stream_t st) {
  RCOT_REQUIRE(out && x && B > 0 && B <= 65535 && n > 0, "axpby: bad arguments");
  dim3 grid(cdiv(n, 256), B);
  axpby_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(out, out_bs, x, x_bs, y, y_bs, a, b, a_vec, a_vec ? 1 : 0, n);
  return check_launch("axpby");
}

extern "C" int rcot_channel_sum(const float* x, int64_t x_bs, float* out, int B, int C, int HW, rcot_stream_t st) {
  RCOT_REQUIRE(x && out && B > 0 && C > 0 && C <= 65535 && HW > 0, "channel_sum: bad arguments");
  long total = (long)B * HW;
  int chunks = (int)((total + 256 * 8 - 1) / (256 * 8));
  if (chunks < 1) chunks = 1;
  if (chunks > 32) chunks = 32;
  dim3 grid(chunks, C);
  channel_sum_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(x, x_bs, out, B, HW);
  return check_launch("channel_sum");
}
