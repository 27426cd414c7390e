// data.cu -- device-side training-patch assembly: what the reference's CPU TrainDataset.__getitem__ does
// per sample (util/dataset_utils.py:215-278) for a whole batch in one launch, from uint8 HWC images resident in HBM:
//   crop_img(base=16) centre crop (util/image_utils.py:59-65) -> random P x P crop -> one of the 8 flip/rot90
//   augmentations (util/image_utils.py:133-163, modes 1..7 drawn by random_augmentation :177-182) -> for the denoise
//   tasks  degraded = uint8(clip(clean + noise * sigma, 0, 255))  (util/degradation_utils.py:21-27, float64 arithmetic
//   like numpy) -> ToTensor (CHW, /255).
// The random draws (image id, crop origin, mode, noise) are INPUTS, so the result is bit-identical to the numpy
// restatement in oracle/data_ref.py for the same draws.  HBM-bound byte shuffling: 3 B in, 24 B out per pixel.
#include "../../include/rcot_b200.h"
#include "common.cuh"

namespace rcot {

__global__ void __launch_bounds__(256)
    make_patches_kernel(const uint8_t* __restrict__ pool, const rcot_patch_desc* __restrict__ desc,
                        const float* __restrict__ noise, float* __restrict__ degraded, float* __restrict__ clean, int B,
                        int P) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= P * P) return;
  const rcot_patch_desc d = desc[b];
  const int i = pix / P, j = pix - i * P;          // output (row, col) AFTER augmentation
  int si, sj;                                      // source (row, col) inside the P x P crop
  switch (d.mode) {                                // np.flipud / np.rot90 compositions, util/image_utils.py:133-163
    case 1: si = P - 1 - i; sj = j; break;                 // flipud
    case 2: si = j; sj = P - 1 - i; break;                 // rot90
    case 3: si = j; sj = i; break;                         // rot90 + flipud
    case 4: si = P - 1 - i; sj = P - 1 - j; break;         // rot180
    case 5: si = i; sj = P - 1 - j; break;                 // rot180 + flipud
    case 6: si = P - 1 - j; sj = i; break;                 // rot270
    case 7: si = P - 1 - j; sj = P - 1 - i; break;         // rot270 + flipud
    default: si = i; sj = j; break;
  }
  // crop_img(base=16): rows [ch/2, H - ch + ch/2), ch = H % 16 (likewise columns); then the random crop origin
  const int ch = d.H % 16, cw = d.W % 16;
  const int y = ch / 2 + d.y0 + si, x = cw / 2 + d.x0 + sj;
  const uint8_t* cp = pool + d.clean_off + ((size_t)y * d.W + x) * 3;
  const size_t o = (size_t)b * 3 * P * P + pix;
  const size_t PP = (size_t)P * P;
  uint8_t c3[3] = {cp[0], cp[1], cp[2]};
  uint8_t g3[3];
  if (d.sigma > 0) {
    // noise is laid out like np.random.randn(P, P, 3) on the AUGMENTED patch
    const float* np_ = noise + ((size_t)b * PP + pix) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double v = (double)c3[c] + (double)np_[c] * (double)d.sigma;
      v = fmin(fmax(v, 0.0), 255.0);
      g3[c] = (uint8_t)v;                          // astype(np.uint8): truncation
    }
  } else {
    const uint8_t* dp = pool + d.deg_off + ((size_t)y * d.W + x) * 3;
    g3[0] = dp[0]; g3[1] = dp[1]; g3[2] = dp[2];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    clean[o + c * PP] = (float)c3[c] / 255.0f;     // ToTensor: uint8 -> float32 / 255
    degraded[o + c * PP] = (float)g3[c] / 255.0f;
  }
}

}  // namespace rcot

extern "C" int rcot_make_patches(const uint8_t* pool, const rcot_patch_desc* desc, const float* noise, float* degraded,
                                 float* clean, int B, int P, rcot_stream_t st) {
  using namespace rcot;
  RCOT_REQUIRE(pool && desc && degraded && clean && B > 0 && P > 0, "make_patches: bad arguments");
  RCOT_REQUIRE(B <= 65535, "make_patches: batch too large");
  dim3 grid(cdiv((long)P * P, 256), B);
  make_patches_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(pool, desc, noise, degraded, clean, B, P);
  return check_launch("make_patches");
}
