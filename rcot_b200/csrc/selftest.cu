// selftest.cu -- bring-up test for the tcgen05 path: D[128 x N] = A[128 x K] * B[N x K]^T
// through the same operand layout, descriptor, pipeline and TMEM read-back code the
// production GEMM kernels use.  Exercised by tests/test_tc_selftest.py on the GPU.
#include "common.cuh"
#include "tc.cuh"

namespace rcot {

template <int TERMS>
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                          float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int STAGES = 2;
  constexpr int A_TILE = op_tile_bytes(128);  // bytes of one bf16 A tile
  constexpr int B_TILE = op_tile_bytes(256);
  constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t cols = tmem_cols_pow2(N);
  if (warp == 0) tmem_alloc(&tmem_base_s, cols);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(128, N);

  const int nchunks = K / KC;
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % STAGES, use = c / STAGES;
    if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
    uint8_t* st = smem + s * STAGE_BYTES;
    uint8_t *a_hi = st, *a_lo = st + A_TILE, *b_hi = st + 2 * A_TILE, *b_lo = st + 2 * A_TILE + B_TILE;
    float v[8];
    for (int k8 = 0; k8 < KC / 8; ++k8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * K + c * KC + k8 * 8 + i];
      op_store8<TERMS>(a_hi, a_lo, tid, k8, v);
    }
    for (int n = tid; n < N; n += 128) {
      for (int k8 = 0; k8 < KC / 8; ++k8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = Bm[(size_t)n * K + c * KC + k8 * 8 + i];
        op_store8<TERMS>(b_hi, b_lo, n, k8, v);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_stage<TERMS>(tmem, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), idesc, c == 0);
      tc_commit(&empty_bar[s]);
    }
  }
  if (tid == 0) tc_commit(&done_bar);
  mbar_wait(&done_bar, 0);
  tc_fence_after();

  const uint32_t lane_base = tmem_lane_base(tmem);
  for (int n0 = 0; n0 < N; n0 += 8) {
    float v[8];
    tmem_ld8(lane_base + n0, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) D[(size_t)tid * N + n0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, cols);
}

}  // namespace rcot

extern "C" int rcot_selftest_tc(const float* A, const float* B, float* D, int N, int K, int terms,
                                cudaStream_t stream) {
  using namespace rcot;
  RCOT_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "selftest: N must be a multiple of 16 in [16,256], got %d", N);
  RCOT_REQUIRE(K >= KC && K % KC == 0, "selftest: K must be a positive multiple of %d, got %d", KC, K);
  RCOT_REQUIRE(terms == 1 || terms == 3, "selftest: terms must be 1 or 3");
  const int smem = 2 * (2 * op_tile_bytes(128) + 2 * op_tile_bytes(256));
  if (terms == 3) {
    cudaFuncSetAttribute(tc_selftest_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_selftest_kernel<3><<<1, 128, smem, stream>>>(A, B, D, N, K);
  } else {
    cudaFuncSetAttribute(tc_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_selftest_kernel<1><<<1, 128, smem, stream>>>(A, B, D, N, K);
  }
  return check_launch("tc_selftest");
}
