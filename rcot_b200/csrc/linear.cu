// linear.cu -- the three fully connected layers at the tail of the potential F_net
// (Net_Restormer.py:496-498,512-520): fc (P^2/2 -> P^2/8), fc1 (-> 64), fc2 (-> 1).
// Batch is at most a few dozen rows, so these are weight-bandwidth-bound (fc.weight is 67 MB at P=128):
// fp32 CUDA-core kernels with exact fp32 accumulation; forward and data gradient share one shared-memory
// tiled GEMM that reads each weight once.  Not GEMM-shaped enough for tcgen05 (M = batch <= 64).
#include "../../include/rcot_b200.h"
#include "common.cuh"

namespace rcot {

// Tiled fp32 GEMM behind the forward and the data gradient:
//     C[b, n] = epi( sum_r A[b, r] * Wv(r, n) ),   A row-major [B, R], C row-major [B, N]
//     NN == false (forward):       Wv(r, n) = W[n * R + r]    (W = [O, K], r = k, n = o)
//     NN == true  (data gradient): Wv(r, n) = W[r * N + n]    (W = [O, K], r = o, n = k)
// CTA = 64 batch rows x 64 columns, 256 threads with a 4 x 4 register tile each; the reduction runs in
// chunks of 32 staged through shared memory (r-major, so both operands are read as float4), the next
// chunk's global loads are issued into registers before the current one is multiplied.  gridDim.y > 1
// splits the reduction; the partial tiles are then added with fp32 atomics into a zeroed C (only used
// without activation / mask; the bias is added by split 0).
constexpr int LT = 64;        // tile edge (batch rows and columns)
constexpr int LRC = 32;       // reduction chunk
constexpr int LLD = LT + 4;   // shared-memory row stride (floats): 16-byte aligned rows

template <bool NN>
__global__ void __launch_bounds__(256)
    linear_gemm_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                       const float* __restrict__ mask, float* __restrict__ Cout, int B, int R, int N, int act,
                       float slope, int r_per_split, int vec_ok) {
  __shared__ __align__(16) float As[LRC][LLD];
  __shared__ __align__(16) float Ws[LRC][LLD];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * LT, b0 = blockIdx.z * LT;
  const int r_begin = blockIdx.y * r_per_split;
  const int r_end = min(R, r_begin + r_per_split);
  // loader roles.  A (and W when r is contiguous): row = tid & 63, eight consecutive r at (tid >> 6) * 8:
  // one 32-byte sector per thread, conflict-free transposed stores.  W with n contiguous: row r = tid >> 3,
  // eight consecutive n at (tid & 7) * 8: coalesced float4 loads and stores.
  const int l_row = tid & 63, l_r8 = (tid >> 6) * 8;
  const int w_r = tid >> 3, w_n8 = (tid & 7) * 8;
  float pa[8], pw[8];
  auto load_chunk = [&](int rc) {
    {
      const int b = b0 + l_row, r = rc + l_r8;
      const float* src = A + (size_t)b * R + r;
      if (vec_ok && b < B && r + 8 <= r_end) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        pa[0] = v0.x; pa[1] = v0.y; pa[2] = v0.z; pa[3] = v0.w;
        pa[4] = v1.x; pa[5] = v1.y; pa[6] = v1.z; pa[7] = v1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pa[i] = (b < B && r + i < r_end) ? __ldg(src + i) : 0.f;
      }
    }
    if (!NN) {
      const int n = n0 + l_row, r = rc + l_r8;
      const float* src = W + (size_t)n * R + r;
      if (vec_ok && n < N && r + 8 <= r_end) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        pw[0] = v0.x; pw[1] = v0.y; pw[2] = v0.z; pw[3] = v0.w;
        pw[4] = v1.x; pw[5] = v1.y; pw[6] = v1.z; pw[7] = v1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pw[i] = (n < N && r + i < r_end) ? __ldg(src + i) : 0.f;
      }
    } else {
      const int r = rc + w_r, n = n0 + w_n8;
      const float* src = W + (size_t)r * N + n;
      if (vec_ok && r < r_end && n + 8 <= N) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        pw[0] = v0.x; pw[1] = v0.y; pw[2] = v0.z; pw[3] = v0.w;
        pw[4] = v1.x; pw[5] = v1.y; pw[6] = v1.z; pw[7] = v1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pw[i] = (r < r_end && n + i < N) ? __ldg(src + i) : 0.f;
      }
    }
  };
  auto store_chunk = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[l_r8 + i][l_row] = pa[i];
    if (!NN) {
#pragma unroll
      for (int i = 0; i < 8; ++i) Ws[l_r8 + i][l_row] = pw[i];
    } else {
      *reinterpret_cast<float4*>(&Ws[w_r][w_n8]) = make_float4(pw[0], pw[1], pw[2], pw[3]);
      *reinterpret_cast<float4*>(&Ws[w_r][w_n8 + 4]) = make_float4(pw[4], pw[5], pw[6], pw[7]);
    }
  };
  const int tb = (tid & 15) * 4, tn = (tid >> 4) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (r_begin < r_end) load_chunk(r_begin);
  for (int rc = r_begin; rc < r_end; rc += LRC) {
    __syncthreads();            // previous chunk fully consumed
    store_chunk();
    __syncthreads();
    if (rc + LRC < r_end) load_chunk(rc + LRC);
#pragma unroll 8
    for (int r = 0; r < LRC; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&As[r][tb]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[r][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
  }
  const bool split = gridDim.y > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + tb + i;
    if (b >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias && blockIdx.y == 0) v += __ldg(bias + n);
      if (split) {
        atomicAdd(Cout + (size_t)b * N + n, v);
      } else {
        if (act) v = v > 0.f ? v : v * slope;
        if (mask) v *= (__ldg(mask + (size_t)b * N + n) > 0.f) ? 1.f : slope;
        Cout[(size_t)b * N + n] = v;
      }
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

// Shared launcher: picks the reduction split so that the grid fills the 148 SMs when the tile grid alone
// does not (only the wide fc layer needs it), zeroing the result first in that case.
template <bool NN>
static int launch_linear(const float* A, const float* W, const float* bias, const float* mask, float* Cout, int B,
                         int R, int N, int act, float slope, cudaStream_t st, const char* what) {
  const int nt = cdiv(N, LT), bt = cdiv(B, LT);
  int S = 1;
  if (!act && !mask) {
    S = cdiv(2 * 148, (long)nt * bt);
    const int maxS = R / (4 * LRC);          // at least four chunks per split
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
  }
  int per = round_up(cdiv(R, S), LRC);
  S = cdiv(R, per);
  RCOT_REQUIRE(bt <= 65535 && S <= 65535, "%s: grid too large", what);
  if (S > 1) {
    cudaError_t e = cudaMemsetAsync(Cout, 0, (size_t)B * N * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("%s: cudaMemsetAsync: %s", what, cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
  }
  const int vec_ok = (R % 4 == 0) && (N % 4 == 0 || !NN) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0);
  dim3 grid(nt, S, bt);
  linear_gemm_kernel<NN><<<grid, 256, 0, st>>>(A, W, bias, mask, Cout, B, R, N, act, slope, per, vec_ok);
  return check_launch(what);
}

extern "C" int rcot_linear_fwd(const float* x, const float* W, const float* bias, const float* mask, float* y, int B,
                               int K, int O, int act, float slope, rcot_stream_t st) {
  RCOT_REQUIRE(x && W && y && B > 0 && K > 0 && O > 0, "linear_fwd: bad arguments");
  return launch_linear<false>(x, W, bias, mask, y, B, K, O, act, slope, (cudaStream_t)st, "linear_fwd");
}

extern "C" int rcot_linear_dgrad(const float* dy, const float* W, const float* mask, float* dx, int B, int K, int O,
                                 float slope, rcot_stream_t st) {
  RCOT_REQUIRE(dy && W && dx && B > 0 && K > 0 && O > 0, "linear_dgrad: bad arguments");
  return launch_linear<true>(dy, W, nullptr, mask, dx, B, O, K, 0, slope, (cudaStream_t)st, "linear_dgrad");
}

extern "C" int rcot_linear_wgrad(const float* dy, const float* x, float* dW, float* db, int B, int K, int O,
                                 rcot_stream_t st) {
  RCOT_REQUIRE(dy && x && dW && B > 0 && K > 0 && O > 0 && O <= 65535, "linear_wgrad: bad arguments");
  dim3 grid(cdiv(K, 256), O);
  linear_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(dy, x, dW, db, B, K, O);
  return check_launch("linear_wgrad");
}
