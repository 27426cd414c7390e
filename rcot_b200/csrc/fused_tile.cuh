// fused_tile.cuh -- geometry and small helpers shared by the one-kernel GDFN forward (gdfn_fused.cu) and the one-kernel
// MDTA phase 1 (mdta_fused.cu): 8 x 16-pixel core tiles with a 1-pixel halo, hidden channels walked in slices of 32.
#pragma once
#include "tc.cuh"

namespace rcot {

constexpr int GF_TH = 8, GF_TW = 16;                 // core tile (pixels)
constexpr int GF_HH = GF_TH + 2, GF_HW = GF_TW + 2;  // halo tile
constexpr int GF_NHP = GF_HH * GF_HW;                // 180 halo pixels
constexpr int GF_RS = 20;                            // shared-memory row stride of a halo row (floats)
constexpr int GF_CS = 228;                           // channel stride (floats) >= 10 * 20 and == 4 (mod 32): the 8 pairs a
                                                     // quarter-warp reads in one 16-byte access fall into 8 distinct bank groups
constexpr int GF_HS = 16;                            // channel pairs per hidden slice (32 channels)

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// 16-byte shared-memory load the compiler may not narrow (a quarter-warp phase of 8 lanes is conflict-free here; the
// 8-byte form's half-warp phase is not).
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

// Named barrier among the 16 stencil warps only (drain and issuer warps never join it).
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

}  // namespace rcot
