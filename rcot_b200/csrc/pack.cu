// pack.cu -- weight pre-packing for the tcgen05 B operand.
// A [N x K] fp32 matrix (any of the strided views described in rcot_b200.h) becomes bf16 hi/lo
// images tiled [pass][k-chunk][term][sub-tile][BN x 32] in the no-swizzle K-major core-matrix
// layout, so a GEMM stage is ONE contiguous cp.async.bulk.  One launch packs a whole table of
// tensors (all weights of a network after an optimizer step).
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "tc.cuh"

namespace rcot {

__global__ void pack_weights_kernel(const rcot_pack_desc* __restrict__ descs) {
  const rcot_pack_desc d = descs[blockIdx.y];
  // plan (must match make_nplan on the host)
  const int nst = (d.N + 255) / 256;
  const int BN = ((d.N + nst - 1) / nst + 15) / 16 * 16;
  const int nk = (d.K + KC - 1) / KC;
  const int Kp = nk * KC;
  const size_t total = (size_t)nst * BN * Kp;
  uint8_t* base = reinterpret_cast<uint8_t*>(d.dst);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(e / Kp), k = (int)(e - (size_t)n * Kp);   // n, k in the padded index space
    float w = 0.f;
    if (n < d.N && k < d.K)
      w = d.src[(size_t)(k / d.R) * d.s_kouter + (size_t)(k % d.R) * (d.s_kinner ? d.s_kinner : 1) + (size_t)n * d.s_n];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    // padded rows of the last pass: n may exceed N but stays inside the padded image
    *reinterpret_cast<__nv_bfloat16*>(base + packed_offset(d.N, d.K, n, k, 0)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + packed_offset(d.N, d.K, n, k, 1)) = lo;
  }
}

}  // namespace rcot

extern "C" size_t rcot_packed_bytes(int N, int K) {
  using namespace rcot;
  NPlan pl = make_nplan(N);
  return (size_t)pl.passes * cdiv(K, KC) * 2 * op_tile_bytes(pl.BN);
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
