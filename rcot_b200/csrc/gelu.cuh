// gelu.cuh -- exact-erf GELU and its derivative (F.gelu default, Net_Restormer.py:83) for the CUDA-core kernels.
#pragma once
#include <cuda_runtime.h>

namespace rcot {

// gelu(a) = a*Phi(a) and gelu'(a) = Phi(a) + a*phi(a) from ONE exponential: Phi through the Abramowitz-Stegun
// 7.1.26 rational form of erfc (absolute error 1.5e-7 in erf, i.e. < 1e-7 in Phi -- fp32 rounding level), whose
// exp(-a^2/2) factor is the same one phi needs.
__device__ __forceinline__ void gelu_pair(float a, float& ge, float& dge) {
  const float x = fabsf(a) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, x, 1.f));
  const float e = __expf(-x * x);
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  const float tail = 0.5f * q * t * e;            // Phi(-|a|)
  const float Phi = a >= 0.f ? 1.f - tail : tail;
  ge = a * Phi;
  dge = fmaf(a * 0.39894228040143268f, e, Phi);
}

__device__ __forceinline__ float gelu_fast(float a) {
  float ge, dge;
  gelu_pair(a, ge, dge);   // the derivative is dead code here
  return ge;
}

}  // namespace rcot
