// pack.cu -- weight pre-packing for the tcgen05 B operand.
// A [N x K] fp32 matrix (any of the strided views described in rcot_b200.h) becomes bf16 hi/lo
// images tiled [pass][k-chunk][term][sub-tile][BN x 32] in the no-swizzle K-major core-matrix
// layout, so a GEMM stage is ONE contiguous cp.async.bulk.  One launch packs a whole table of
// tensors (all weights of a network after an optimizer step).
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "tc.cuh"

namespace rcot {

__device__ __forceinline__ void pack_one(const rcot_pack_desc& d, size_t e) {
  // plan (must match make_nplan on the host)
  int BN, nsub_total, NSUB, passes;
  if (d.N <= 256) {
    nsub_total = 1;
    BN = (d.N + 15) / 16 * 16;
    NSUB = 1;
    passes = 1;
  } else {
    nsub_total = (d.N + 255) / 256;
    BN = ((d.N + nsub_total - 1) / nsub_total + 15) / 16 * 16;
    NSUB = 2;
    passes = (nsub_total + 1) / 2;
  }
  const int nk = (d.K + KC - 1) / KC;
  const int Kp = nk * KC;
  const size_t total = (size_t)passes * NSUB * BN * Kp;
  if (e >= total) return;
  const int n = (int)(e / Kp), k = (int)(e - (size_t)n * Kp);
  float w = 0.f;
  if (n < d.N && k < d.K) w = d.src[(size_t)(k / d.R) * d.s_kouter + (k % d.R) + (size_t)n * d.s_n];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const int si = n / BN, np = n - si * BN;
  const int pass = si / NSUB, sub = si - pass * NSUB;
  const int kc = k / KC, kp = k - kc * KC;
  const size_t tile = (size_t)BN * KC * 2;
  uint8_t* base = reinterpret_cast<uint8_t*>(d.dst);
  const size_t off_hi = ((((size_t)pass * nk + kc) * 2 + 0) * NSUB + sub) * tile + op_offset(np, kp);
  const size_t off_lo = ((((size_t)pass * nk + kc) * 2 + 1) * NSUB + sub) * tile + op_offset(np, kp);
  *reinterpret_cast<__nv_bfloat16*>(base + off_hi) = hi;
  *reinterpret_cast<__nv_bfloat16*>(base + off_lo) = lo;
}

__global__ void pack_weights_kernel(const rcot_pack_desc* __restrict__ descs) {
  const rcot_pack_desc d = descs[blockIdx.y];
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;; e += (size_t)gridDim.x * blockDim.x) {
    // bound check inside pack_one needs the plan; recompute the cheap upper bound here
    const int nk = (d.K + KC - 1) / KC;
    size_t rows;
    if (d.N <= 256) rows = (d.N + 15) / 16 * 16;
    else {
      const int nst = (d.N + 255) / 256;
      rows = (size_t)((nst + 1) / 2) * 2 * (((d.N + nst - 1) / nst + 15) / 16 * 16);
    }
    if (e >= rows * nk * KC) break;
    pack_one(d, e);
  }
}

}  // namespace rcot

extern "C" size_t rcot_packed_bytes(int N, int K) {
  using namespace rcot;
  NPlan pl = make_nplan(N);
  return (size_t)pl.passes * pl.NSUB * pl.BN * cdiv(K, KC) * KC * 2 * 2;
}

extern "C" int rcot_pack_weights(const rcot_pack_desc* descs, int n, size_t max_elems, rcot_stream_t stream_) {
  using namespace rcot;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RCOT_REQUIRE(descs != nullptr && n > 0, "pack_weights: empty table");
  RCOT_REQUIRE(n <= 65535, "pack_weights: too many tensors in one table (%d)", n);
  long blocks = (long)((max_elems + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)n);
  pack_weights_kernel<<<grid, 256, 0, stream>>>(descs);
  return check_launch("pack_weights");
}
