// tc.cuh -- sm_100a primitives: mbarrier, 1-D bulk TMA copy, TMEM allocation,
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) and tcgen05.ld, plus the
// no-swizzle K-major operand layout every GEMM-shaped kernel in this library uses.
//
// Operand layout (shared memory, "core matrix" = 8 rows x 16 bytes, stored as 128
// contiguous bytes): element (row r, k) of a [rows x KC] bf16 tile lives at
//     (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2        bytes
// with LBO = 160 (core matrix + 32 B pad, see OP_LBO) and SBO = (KC/8)*LBO.
// Threads produce the operands (fp32 -> bf16 hi [+ lo]), so no TMA swizzle mode has
// to be matched; weights are pre-packed in this layout in HBM and brought in with
// cp.async.bulk (UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace rcot {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Pure polling wait (no suspension), bounded like mbar_wait.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  if (mbar_test_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_test_wait(bar, parity)) {
    if (clock64() - t0 > 3000000000LL) {
      printf("rcot: mbarrier spin wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// Bounded wait: a barrier that never completes traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 3000000000LL) {
      printf("rcot: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completes on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Make generic-proxy shared-memory writes visible to the async proxy (tensor core / TMA).
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ TMEM
// One full warp calls alloc/dealloc. ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__host__ __device__ inline uint32_t tmem_cols_pow2(uint32_t n) {
  uint32_t c = 32;
  while (c < n) c <<= 1;
  return c;
}

// True in exactly one lane of a fully active warp (elect.sync): unlike `lane == 0` the compiler KNOWS the guarded code
// runs in a single thread, so uniform-register operands (tcgen05.mma descriptors) need no waterfall loop per instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE, K-major (bit layout as documented for
// sm_100 UMMA: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48)).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor: D=f32, A=B=bf16, both K-major, M x N tile.
__host__ __device__ inline uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand RESIDENT IN TENSOR MEMORY ("TS" form): row m of A in TMEM lane m, two bf16 K-elements per
// 32-bit column (even k in the low half) -- verified by rcot_selftest_tmem_a.  One k16 step reads 8 columns.
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM, 4 consecutive columns of this thread's lane (warp-collective; pair with tmem_st_wait()).
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Arrive on `bar` when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------ TMEM -> registers
// 32x32b: thread t of warp w reads lane 32*(w%4)+t, consecutive columns.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 columns, no wait: pair with tmem_ld_wait() before the registers are read.
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tmem_lane_base(uint32_t tmem_base) {
  // lane field is bits [31:16]; each warp may only touch its own 32-lane quarter.
  return tmem_base + ((((threadIdx.x >> 5) & 3u) * 32u) << 16);
}

// ------------------------------------------------------------------ operand packing
constexpr int KC = 32;                 // K elements per pipeline stage
// LBO = 160 (a 128-byte core matrix + 32 bytes of padding): with 128 the four K-adjacent core matrices of
// a row group start in the same shared-memory banks, and a quarter-warp that stores {2 rows x 4 k-groups}
// (the coalescing-friendly mapping of the pixel-as-K kernel) would hit 4-way bank conflicts.
constexpr uint32_t OP_LBO = 160;                 // bytes between core matrices adjacent in K
constexpr uint32_t OP_SBO = (KC / 8) * OP_LBO;   // bytes between 8-row groups
__host__ __device__ constexpr uint32_t op_tile_bytes(int rows) { return (uint32_t)(rows / 8) * OP_SBO; }

__device__ __host__ inline uint32_t op_offset(int row, int k) {  // k in [0,KC)
  return (uint32_t)(row >> 3) * OP_SBO + (uint32_t)(k >> 3) * OP_LBO + (uint32_t)(row & 7) * 16u +
         (uint32_t)(k & 7) * 2u;
}

// Byte offset of element (n, k), term (0 = hi, 1 = lo), inside the packed B-operand image of an
// [N x K] matrix: tiles ordered [pass][k-chunk][term][sub-tile][BN x 32] (see pack.cu / make_nplan).
__device__ __host__ inline size_t packed_offset(int N, int K, int n, int k, int term) {
  const int nst = (N + 255) / 256;
  const int BN = ((N + nst - 1) / nst + 15) / 16 * 16;
  const int nk = (K + KC - 1) / KC;
  const int pass = n / BN, np = n - pass * BN;
  const int kc = k / KC, kp = k - kc * KC;
  const size_t tile = op_tile_bytes(BN);
  return (((size_t)pass * nk + kc) * 2 + term) * tile + op_offset(np, kp);
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// Split 8 fp32 values into bf16 hi and lo (x ~= hi + lo, |err| <= 2^-17 |x|).  Pairs are converted
// with the packed cvt.rn.bf16x2.f32 (F2FP, full rate); hi is widened back with shifts/masks.
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);   // .x = low half = v[2i]
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hp);
    h[i] = hb;
    const float r0 = v[2 * i] - __uint_as_float(hb << 16);
    const float r1 = v[2 * i + 1] - __uint_as_float(hb & 0xffff0000u);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(r0, r1);
    l[i] = *reinterpret_cast<const uint32_t*>(&lp);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// Store 8 consecutive-k values of one operand row (k8 = k/8 within the stage).
template <int TERMS>
__device__ __forceinline__ void op_store8(uint8_t* hi_tile, uint8_t* lo_tile, int row, int k8, const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  uint32_t off = (uint32_t)(row >> 3) * OP_SBO + (uint32_t)k8 * OP_LBO + (uint32_t)(row & 7) * 16u;
  *reinterpret_cast<uint4*>(hi_tile + off) = hi;
  if (TERMS > 1) *reinterpret_cast<uint4*>(lo_tile + off) = lo;
}

// Store 8 consecutive-k bf16 values (already exact) of one operand row.
__device__ __forceinline__ void op_store8_bf16(uint8_t* hi_tile, int row, int k8, const uint32_t (&pk)[4]) {
  const uint32_t off = (uint32_t)(row >> 3) * OP_SBO + (uint32_t)k8 * OP_LBO + (uint32_t)(row & 7) * 16u;
  *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// Issue the MMAs of one K stage (KC = 2 k16 steps) for one [128 x BN] accumulator.
// TERMS==3: hi*hi + lo*hi + hi*lo (fp32-class accuracy); TERMS==1: hi*hi (bf16 compute).
// ALO = false: the A operand is exact in bf16 (bf16-stored activations): its lo term does not exist.
// BLO = false: likewise for B.
template <int TERMS, bool ALO = true, bool BLO = true>
__device__ __forceinline__ void issue_stage(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                            uint32_t b_lo, uint32_t idesc, bool first_stage) {
#pragma unroll
  for (int s = 0; s < KC / 16; ++s) {
    uint32_t ko = s * 2 * OP_LBO;
    uint64_t dah = make_sdesc(a_hi + ko, OP_LBO, OP_SBO);
    uint64_t dbh = make_sdesc(b_hi + ko, OP_LBO, OP_SBO);
    tc_mma_bf16(tmem_d, dah, dbh, idesc, (first_stage && s == 0) ? 0u : 1u);
    if (TERMS > 1) {
      if (ALO) {
        uint64_t dal = make_sdesc(a_lo + ko, OP_LBO, OP_SBO);
        tc_mma_bf16(tmem_d, dal, dbh, idesc, 1u);
      }
      if (BLO) {
        uint64_t dbl = make_sdesc(b_lo + ko, OP_LBO, OP_SBO);
        tc_mma_bf16(tmem_d, dah, dbl, idesc, 1u);
      }
    }
  }
}

}  // namespace rcot
