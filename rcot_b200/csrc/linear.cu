// linear.cu -- the three fully connected layers at the tail of the potential F_net
// (Net_Restormer.py:496-498,512-520): fc (P^2/2 -> P^2/8), fc1 (-> 64), fc2 (-> 1).
// Batch is at most a few dozen rows, so these are weight-bandwidth-bound matrix-vector bundles:
// plain fp32 CUDA-core kernels (exact fp32 accumulation), each weight read once per pass from
// HBM/L2, batch rows held in registers.  Not GEMM-shaped enough for tcgen05 (M = batch <= 64).
#include "../../include/rcot_b200.h"
#include "common.cuh"

namespace rcot {

constexpr int LB = 8;  // batch rows per register tile

__device__ __forceinline__ float warp_sum_l(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// y[b,o] = act( sum_k x[b,k] W[o,k] + bias[o] ) * maskfactor.  One warp per output feature.
__global__ void __launch_bounds__(256)
    linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                      const float* __restrict__ mask, float* __restrict__ y, int B, int K, int O, int act,
                      float slope) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= O) return;
  const float* w = W + (size_t)o * K;
  for (int b0 = 0; b0 < B; b0 += LB) {
    float acc[LB];
#pragma unroll
    for (int i = 0; i < LB; ++i) acc[i] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(w + k);
#pragma unroll
      for (int i = 0; i < LB; ++i)
        if (b0 + i < B) acc[i] = fmaf(__ldg(x + (size_t)(b0 + i) * K + k), wv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const float s = warp_sum_l(acc[i]);
      if (lane == 0 && b0 + i < B) {
        float v = s + (bias ? __ldg(bias + o) : 0.f);
        if (act) v = v > 0.f ? v : v * slope;
        if (mask) v *= (__ldg(mask + (size_t)(b0 + i) * O + o) > 0.f) ? 1.f : slope;
        y[(size_t)(b0 + i) * O + o] = v;
      }
    }
  }
}

// dx[b,k] = ( sum_o dy[b,o] W[o,k] ) * maskfactor(mask[b,k]).  One thread per input feature k.
__global__ void __launch_bounds__(256)
    linear_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ W, const float* __restrict__ mask,
                        float* __restrict__ dx, int B, int K, int O, float slope) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = blockIdx.y * LB;
  if (k >= K) return;
  float acc[LB];
#pragma unroll
  for (int i = 0; i < LB; ++i) acc[i] = 0.f;
  for (int o = 0; o < O; ++o) {
    const float wv = __ldg(W + (size_t)o * K + k);
#pragma unroll
    for (int i = 0; i < LB; ++i)
      if (b0 + i < B) acc[i] = fmaf(__ldg(dy + (size_t)(b0 + i) * O + o), wv, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < LB; ++i) {
    if (b0 + i < B) {
      float v = acc[i];
      if (mask) v *= (__ldg(mask + (size_t)(b0 + i) * K + k) > 0.f) ? 1.f : slope;
      dx[(size_t)(b0 + i) * K + k] = v;
    }
  }
}

// dW[o,k] += sum_b dy[b,o] x[b,k] ;  db[o] += sum_b dy[b,o]
__global__ void __launch_bounds__(256)
    linear_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dW,
                        float* __restrict__ db, int B, int K, int O) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y;
  if (k >= K) return;
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = __ldg(dy + (size_t)b * O + o);
    acc = fmaf(d, __ldg(x + (size_t)b * K + k), acc);
    accb += d;
  }
  dW[(size_t)o * K + k] += acc;
  if (db && k == 0) db[o] += accb;
}

}  // namespace rcot

using namespace rcot;

extern "C" int rcot_linear_fwd(const float* x, const float* W, const float* bias, const float* mask, float* y, int B,
                               int K, int O, int act, float slope, rcot_stream_t st) {
  RCOT_REQUIRE(x && W && y && B > 0 && K > 0 && O > 0, "linear_fwd: bad arguments");
  linear_fwd_kernel<<<cdiv(O, 8), 256, 0, (cudaStream_t)st>>>(x, W, bias, mask, y, B, K, O, act, slope);
  return check_launch("linear_fwd");
}

extern "C" int rcot_linear_dgrad(const float* dy, const float* W, const float* mask, float* dx, int B, int K, int O,
                                 float slope, rcot_stream_t st) {
  RCOT_REQUIRE(dy && W && dx && B > 0 && K > 0 && O > 0, "linear_dgrad: bad arguments");
  dim3 grid(cdiv(K, 256), cdiv(B, LB));
  linear_dgrad_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(dy, W, mask, dx, B, K, O, slope);
  return check_launch("linear_dgrad");
}

extern "C" int rcot_linear_wgrad(const float* dy, const float* x, float* dW, float* db, int B, int K, int O,
                                 rcot_stream_t st) {
  RCOT_REQUIRE(dy && x && dW && B > 0 && K > 0 && O > 0 && O <= 65535, "linear_wgrad: bad arguments");
  dim3 grid(cdiv(K, 256), O);
  linear_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(dy, x, dW, db, B, K, O);
  return check_launch("linear_wgrad");
}
