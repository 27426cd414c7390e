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

// ------------------------------------------------------------------ A operand from TMEM (tcgen05.mma "TS" form)
// Bring-up for keeping a resident A operand (e.g. LN(x) of a tile) in tensor memory instead of shared memory: row m of A
// lives in TMEM lane m, two bf16 K-elements per 32-bit column (variant 0: even k in the low half), written with
// tcgen05.st by the thread that owns the lane; each k16 step of the MMA reads 8 columns.  Single bf16 term.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__global__ void __launch_bounds__(128) tc_selftest_tmema_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                float* __restrict__ D, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_col = 256;                          // A: columns [256, 256 + K/2); D: columns [0, N)
  const uint32_t lane_addr = tmem_lane_base(tmem) + a_col;
  const int nchunks = K / KC;
  const uint32_t b_tile = op_tile_bytes(256);
  auto store_a = [&]() {                               // A -> TMEM: this thread owns lane tid
    for (int c8 = 0; c8 < K / 16; ++c8) {              // 8 columns = 16 K-elements per store
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float e0 = A[(size_t)tid * K + c8 * 16 + 2 * i], e1 = A[(size_t)tid * K + c8 * 16 + 2 * i + 1];
        const __nv_bfloat162 pk = (variant & 1) == 0 ? __floats2bfloat162_rn(e0, e1) : __floats2bfloat162_rn(e1, e0);
        r[i] = *reinterpret_cast<const uint32_t*>(&pk);
      }
      tmem_st8(lane_addr + c8 * 8, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  };
  auto store_b = [&]() {                               // B -> shared (hi term only), chunk layout of the production kernels
    float v[8];
    for (int c = 0; c < nchunks; ++c)
      for (int n = tid; n < N; n += 128)
        for (int k8 = 0; k8 < KC / 8; ++k8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = Bm[(size_t)n * K + c * KC + k8 * 8 + i];
          op_store8<1>(smem + c * b_tile, smem + c * b_tile, n, k8, v);
        }
  };
  if (variant & 2) {
    store_b();
    __syncthreads();
    store_a();
  } else {
    store_a();
    store_b();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, N);
    for (int c = 0; c < nchunks; ++c)
      for (int s2 = 0; s2 < KC / 16; ++s2) {
        const uint64_t db = make_sdesc(smem_u32(smem + c * b_tile) + s2 * 2 * OP_LBO, OP_LBO, OP_SBO);
        tc_mma_bf16_ts(tmem, tmem + a_col + (c * (KC / 16) + s2) * 8, db, idesc, (c | s2) ? 1u : 0u);
      }
    tc_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  const uint32_t lane_base = tmem_lane_base(tmem);
  for (int n0 = 0; n0 < N; n0 += 8) {
    float o[8];
    tmem_ld8(lane_base + n0, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) D[(size_t)tid * N + n0 + i] = o[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace rcot

extern "C" int rcot_selftest_tmem_a(const float* A, const float* B, float* D, int N, int K, int variant,
                                    cudaStream_t stream) {
  using namespace rcot;
  RCOT_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "selftest_tmem_a: N must be a multiple of 16 in [16,256], got %d", N);
  RCOT_REQUIRE(K >= KC && K % KC == 0 && K <= 256, "selftest_tmem_a: K must be a multiple of %d up to 256, got %d", KC, K);
  const int smem = (K / KC) * op_tile_bytes(256);
  cudaFuncSetAttribute(tc_selftest_tmema_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tc_selftest_tmema_kernel<<<1, 128, smem, stream>>>(A, B, D, N, K, variant);
  return check_launch("tc_selftest_tmem_a");
}

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
