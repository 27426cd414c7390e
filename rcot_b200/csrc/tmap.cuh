// tmap.cuh -- host-side CUtensorMap construction for fp32 NCHW activations (shared by the TMA-staged GEMM variants).
// The encoder comes from the driver through the runtime (cudaGetDriverEntryPoint): no link dependency on libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

namespace rcot {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// nullptr when the driver does not export cuTensorMapEncodeTiled: callers then stay on their register-staged kernels
// (still sm_100a code -- there is no other fallback).
inline EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn) encode = reinterpret_cast<EncodeTiledFn>(fn);
    (void)cudaGetLastError();
  }
  return encode;
}

// Map of an activation [B, C, HW] (per-image block contiguous, batch stride bs elements) with a box of
// box_px pixels x box_ch channels x 1 image, no swizzle: the box lands in shared memory as [channel][pixel].
// Channels / pixels of a box that fall outside the tensor are filled with zeros.
inline int make_act_map(CUtensorMap* tm, const void* base, int64_t bs, int C, long HW, int B, int box_px, int box_ch,
                        const char* who, int bf16 = 0) {
  EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) {
    set_error("%s: cuTensorMapEncodeTiled is not available from this driver", who);
    return RCOT_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
  const size_t esz = bf16 ? 2 : sizeof(float);
  const cuuint64_t strides[2] = {(cuuint64_t)HW * esz, (cuuint64_t)bs * esz};
  const cuuint32_t box[3] = {(cuuint32_t)box_px, (cuuint32_t)box_ch, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                            const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d) for C=%d HW=%ld B=%d bs=%lld box=%dx%d", who, (int)r, C, HW, B,
              (long long)bs, box_px, box_ch);
    return RCOT_ERR_CUDA;
  }
  return RCOT_OK;
}

// One 3-D tensor copy global -> shared (coordinates: pixel, channel, image), completing on `bar`.
__device__ __forceinline__ void tensor_g2s_3d(void* dst_smem, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z),
      "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
      : "memory");
}

}  // namespace rcot
