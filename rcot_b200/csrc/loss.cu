// loss.cu -- the objective-side kernels of one adversarial iteration (reference trainer.py:262-346):
//   * transport cost of the T-sub step: res = degraded - T(x); RMSE over the whole batch; the
//     Fourier-residual penalty per sample (trainer.py:323-332: `mean(|F|^2)**1/2` parses as
//     mean(|F|^2)/2 for de_id < 3, mean(|F|) otherwise; SUM over the batch); optional paired L1;
//     and the assembled gradient dL/dT(x).  The 2-D FFT of each [P x P] channel plane runs in shared
//     memory: forward decimation-in-frequency, pointwise F/|F|, inverse decimation-in-time -- the
//     bit-reversed orders of the two cancel, so no permutation pass exists.
//   * gradient-penalty coefficients (trainer.py:300-305)
//   * RMSprop / Adam on flat parameter buffers (torch.optim defaults, trainer.py:121-126)
#include "../../include/rcot_b200.h"
#include "common.cuh"

namespace rcot {

__device__ __forceinline__ float warp_sum_c(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum, result valid in thread 0. `red` is a 32-float shared scratch.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum_c(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_sum_c(v);
  }
  return v;
}

// One radix-2 stage over `nlines` lines of length P stored with element stride `es` and line stride `ls`.
// DIF (forward, twiddle e^{-2 pi i j/(2m)}) when inverse == 0, DIT (inverse, conjugate twiddle) otherwise.
__device__ __forceinline__ void fft_stage(float2* s, const float2* tw, int P, int m, int es, int ls, int inverse) {
  const int half = P >> 1;
  const int tstep = half / m;  // twiddle index stride: w_P^(j * P/(2m))
  for (int t = threadIdx.x; t < P * half; t += blockDim.x) {
    const int line = t / half, bf = t - line * half;
    const int grp = bf / m, j = bf - grp * m;
    const int i0 = grp * 2 * m + j;
    float2* p0 = s + line * ls + i0 * es;
    float2* p1 = p0 + m * es;
    float2 w = tw[j * tstep];
    const float2 a = *p0;
    float2 b = *p1;
    if (!inverse) {
      const float2 d = make_float2(a.x - b.x, a.y - b.y);
      *p0 = make_float2(a.x + b.x, a.y + b.y);
      *p1 = make_float2(d.x * w.x - d.y * w.y, d.x * w.y + d.y * w.x);
    } else {
      w.y = -w.y;
      const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      *p0 = make_float2(a.x + bw.x, a.y + bw.y);
      *p1 = make_float2(a.x - bw.x, a.y - bw.y);
    }
  }
  __syncthreads();
}

// Stage 1: one CTA per (channel, image).  acc[0] += sum res^2 (all samples), acc[1] += fourier term,
// acc[2] += sum |out - target|.  gfou[b,c,:,:] = d(fourier_b)/d(res).
__global__ void __launch_bounds__(512)
    cost_stage1_kernel(const float* __restrict__ out, const float* __restrict__ degraded,
                       const float* __restrict__ target, const int64_t* __restrict__ de_id, float* __restrict__ gfou,
                       float* __restrict__ acc, int P) {
  extern __shared__ float2 sm2[];
  const int LS = P + 1;  // padded line stride (float2 units): conflict-free column passes
  float2* s = sm2;
  float2* tw = sm2 + (size_t)P * LS;
  __shared__ float red[32];
  const int ch = blockIdx.x, b = blockIdx.y;
  const size_t base = ((size_t)b * 3 + ch) * P * P;
  const bool l2branch = de_id[b] < 3;
  float s_res2 = 0.f, s_l1 = 0.f;
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) {
    const float o = __ldg(out + base + e);
    const float r = __ldg(degraded + base + e) - o;
    s_res2 = fmaf(r, r, s_res2);
    if (target) s_l1 += fabsf(o - __ldg(target + base + e));
    const int y = e / P, x = e - y * P;
    s[y * LS + x] = make_float2(r, 0.f);
    if (l2branch) gfou[base + e] = r * (1.f / 3.f);  // d/dres of sum(res^2)/6 ... per sample (Parseval)
  }
  for (int k = threadIdx.x; k < P / 2; k += blockDim.x) {
    float sn, cs;
    sincospif(-2.f * (float)k / (float)P, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  const float tot_res2 = block_sum(s_res2, red);
  const float tot_l1 = block_sum(s_l1, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, tot_res2);
    if (target) atomicAdd(acc + 2, tot_l1);
    if (l2branch) atomicAdd(acc + 1, tot_res2 * (1.f / 6.f));
  }
  if (l2branch) return;  // uniform per CTA
  __syncthreads();
  // forward: rows then columns, decimation in frequency (natural in, bit-reversed out)
  for (int m = P / 2; m >= 1; m >>= 1) fft_stage(s, tw, P, m, 1, LS, 0);
  for (int m = P / 2; m >= 1; m >>= 1) fft_stage(s, tw, P, m, LS, 1, 0);
  float s_abs = 0.f;
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) {
    const int y = e / P, x = e - y * P;
    float2 v = s[y * LS + x];
    const float mag = sqrtf(v.x * v.x + v.y * v.y);
    s_abs += mag;
    const float inv = mag > 0.f ? 1.f / mag : 0.f;  // torch: d|z|/dz = 0 at z = 0
    s[y * LS + x] = make_float2(v.x * inv, v.y * inv);
  }
  const float tot_abs = block_sum(s_abs, red);
  if (threadIdx.x == 0) atomicAdd(acc + 1, tot_abs / (3.f * P * P));
  __syncthreads();
  // inverse (unnormalised): columns then rows, decimation in time (bit-reversed in, natural out)
  for (int m = 1; m <= P / 2; m <<= 1) fft_stage(s, tw, P, m, LS, 1, 1);
  for (int m = 1; m <= P / 2; m <<= 1) fft_stage(s, tw, P, m, 1, LS, 1);
  const float scale = 1.f / (3.f * P * P);
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) {
    const int y = e / P, x = e - y * P;
    gfou[base + e] = s[y * LS + x].x * scale;
  }
}

// Stage 2: dL/dout = dF - sigma*res/(N*rmse) - sigma*gfou + Sigma*sign(out-target)/N
// with N = n_global elements and rmse = sqrt(acc[0]/N) (acc[0] already all-reduced for data parallel runs).
__global__ void __launch_bounds__(256)
    cost_stage2_kernel(const float* __restrict__ out, const float* __restrict__ degraded,
                       const float* __restrict__ target, const float* __restrict__ gfou, const float* __restrict__ dF,
                       const float* __restrict__ acc, float* __restrict__ dout, float sigma, float Sigma,
                       float n_global, long n) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float rmse = sqrtf(acc[0] / n_global);
  const float o = __ldg(out + e);
  const float r = __ldg(degraded + e) - o;
  float g = dF ? __ldg(dF + e) : 0.f;
  g -= sigma * (r / (n_global * rmse) + __ldg(gfou + e));
  if (target) {
    const float d = o - __ldg(target + e);
    g += Sigma / n_global * ((d > 0.f) - (d < 0.f));
  }
  dout[e] = g;
}

// per-sample sum of squares: out[b] = sum_e x[b,e]^2
__global__ void __launch_bounds__(256) sample_sumsq_kernel(const float* __restrict__ x, float* __restrict__ out, long n) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  float s = 0.f;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const float v = __ldg(x + (size_t)b * n + e);
    s = fmaf(v, v, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out + b, s);
}

// gradient penalty: coef[b] = 10 * (2/Bglobal) * (|g_b| - 1)/|g_b| ; loss += 10/Bglobal * sum_b (|g_b|-1)^2
__global__ void gp_coef_kernel(const float* __restrict__ sumsq, float* __restrict__ coef, float* __restrict__ loss,
                               int B, float b_global) {
  float l = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float nrm = sqrtf(sumsq[b]);
    coef[b] = nrm > 0.f ? 20.f / b_global * (nrm - 1.f) / nrm : 0.f;
    l += 10.f / b_global * (nrm - 1.f) * (nrm - 1.f);
  }
  l = warp_sum_c(l);
  if (threadIdx.x == 0) atomicAdd(loss, l);
}

// out[0] += scale * sum_i w_i x[i]   with w_i = (i < n_neg ? -1 : +1)   (critic means)
__global__ void signed_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int n_neg, float scale) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (i < n_neg ? -1.f : 1.f) * x[i];
  s = warp_sum_c(s);
  if (threadIdx.x == 0) atomicAdd(out, s * scale);
}

__global__ void __launch_bounds__(256)
    rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq, long n, float lr,
                   float alpha, float eps, float gscale, const float* __restrict__ hyper) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (hyper) lr *= __ldg(hyper);   // device-resident learning rate (CUDA-graph replays), lr = multiplier
  const float gv = g[i] * gscale;
  const float s = alpha * sq[i] + (1.f - alpha) * gv * gv;
  sq[i] = s;
  p[i] -= lr * gv / (sqrtf(s) + eps);
}

__global__ void __launch_bounds__(256)
    adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                long n, float lr, float b1, float b2, float eps, float bc1, float bc2, float gscale,
                const float* __restrict__ hyper) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (hyper) {   // device-resident (lr, 1-b1^t, 1-b2^t) for CUDA-graph replays
    lr *= __ldg(hyper);
    bc1 = __ldg(hyper + 1);
    bc2 = __ldg(hyper + 2);
  }
  const float gv = g[i] * gscale;
  const float mm = b1 * m[i] + (1.f - b1) * gv;
  const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
  m[i] = mm;
  v[i] = vv;
  p[i] -= (lr / bc1) * mm / (sqrtf(vv) / sqrtf(bc2) + eps);
}

}  // namespace rcot

using namespace rcot;

extern "C" int rcot_cost_stage1(const float* out, const float* degraded, const float* target, const int64_t* de_id,
                                float* gfou, float* acc, int B, int P, rcot_stream_t st) {
  RCOT_REQUIRE(out && degraded && de_id && gfou && acc && B > 0 && B <= 65535, "cost_stage1: bad arguments");
  RCOT_REQUIRE(P >= 8 && P <= 128 && (P & (P - 1)) == 0, "cost_stage1: patch size must be a power of two in [8,128], got %d", P);
  const size_t smem = ((size_t)P * (P + 1) + P / 2) * sizeof(float2);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cost_stage1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
    if (e != cudaSuccess) {
      set_error("cost_stage1: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(3, B);
  cost_stage1_kernel<<<grid, 512, smem, (cudaStream_t)st>>>(out, degraded, target, de_id, gfou, acc, P);
  return check_launch("cost_stage1");
}

extern "C" int rcot_cost_stage2(const float* out, const float* degraded, const float* target, const float* gfou,
                                const float* dF, const float* acc, float* dout, float sigma, float Sigma,
                                double n_global, int64_t n, rcot_stream_t st) {
  RCOT_REQUIRE(out && degraded && gfou && acc && dout && n > 0 && n_global > 0, "cost_stage2: bad arguments");
  cost_stage2_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)st>>>(out, degraded, target, gfou, dF, acc, dout, sigma,
                                                                 Sigma, (float)n_global, n);
  return check_launch("cost_stage2");
}

extern "C" int rcot_sample_sumsq(const float* x, float* out, int B, int64_t n, rcot_stream_t st) {
  RCOT_REQUIRE(x && out && B > 0 && B <= 65535 && n > 0, "sample_sumsq: bad arguments");
  int chunks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (chunks > 64) chunks = 64;
  dim3 grid(chunks, B);
  sample_sumsq_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(x, out, n);
  return check_launch("sample_sumsq");
}

extern "C" int rcot_gp_coef(const float* sumsq, float* coef, float* loss, int B, int B_global, rcot_stream_t st) {
  RCOT_REQUIRE(sumsq && coef && loss && B > 0 && B_global > 0, "gp_coef: bad arguments");
  gp_coef_kernel<<<1, 32, 0, (cudaStream_t)st>>>(sumsq, coef, loss, B, (float)B_global);
  return check_launch("gp_coef");
}

extern "C" int rcot_signed_sum(const float* x, float* out, int n, int n_neg, float scale, rcot_stream_t st) {
  RCOT_REQUIRE(x && out && n > 0, "signed_sum: bad arguments");
  signed_sum_kernel<<<1, 32, 0, (cudaStream_t)st>>>(x, out, n, n_neg, scale);
  return check_launch("signed_sum");
}

extern "C" int rcot_rmsprop(float* p, const float* g, float* sq, int64_t n, float lr, float alpha, float eps,
                            float gscale, rcot_stream_t st) {
  RCOT_REQUIRE(p && g && sq && n > 0, "rmsprop: bad arguments");
  rmsprop_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)st>>>(p, g, sq, n, lr, alpha, eps, gscale, nullptr);
  return check_launch("rmsprop");
}

extern "C" int rcot_rmsprop_h(float* p, const float* g, float* sq, int64_t n, const float* hyper, float lr_mult,
                              float alpha, float eps, float gscale, rcot_stream_t st) {
  RCOT_REQUIRE(p && g && sq && hyper && n > 0, "rmsprop_h: bad arguments");
  rmsprop_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)st>>>(p, g, sq, n, lr_mult, alpha, eps, gscale, hyper);
  return check_launch("rmsprop_h");
}

extern "C" int rcot_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
                         float eps, int step, float gscale, rcot_stream_t st) {
  RCOT_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adam: bad arguments");
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  adam_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)st>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, bc2, gscale, nullptr);
  return check_launch("adam");
}

extern "C" int rcot_adam_h(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float lr_mult,
                           float b1, float b2, float eps, float gscale, rcot_stream_t st) {
  RCOT_REQUIRE(p && g && m && v && hyper && n > 0, "adam_h: bad arguments");
  adam_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)st>>>(p, g, m, v, n, lr_mult, b1, b2, eps, 1.f, 1.f, gscale, hyper);
  return check_launch("adam_h");
}
