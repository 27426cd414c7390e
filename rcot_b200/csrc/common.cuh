// common.cuh -- error plumbing and small helpers shared by every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define RCOT_OK 0
#define RCOT_ERR_ARG (-1)
#define RCOT_ERR_CUDA (-2)
#define RCOT_ERR_ARCH (-3)

namespace rcot {

void set_error(const char* fmt, ...);  // api.cu
int check_launch(const char* what);    // api.cu: cudaPeekAtLastError -> error code

#define RCOT_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      rcot::set_error(__VA_ARGS__);    \
      return RCOT_ERR_ARG;             \
    }                                  \
  } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
static inline int round_up(int a, int b) { return ((a + b - 1) / b) * b; }

// Tiling of the N (output-channel) dimension of a pixel-as-M GEMM: `passes` tiles of BN <= 256
// columns each (two TMEM accumulator buffers of BN columns per CTA); tiles along N are separate
// work items of the persistent kernel. NSUB is kept (= 1) so the packed layout formula stays generic.
struct NPlan {
  int N, BN, nsub_total, NSUB, passes;
};
static inline NPlan make_nplan(int N) {
  NPlan p;
  p.N = N;
  p.nsub_total = cdiv(N, 256);
  p.BN = round_up(cdiv(N, p.nsub_total), 16);
  p.NSUB = 1;
  p.passes = p.nsub_total;
  return p;
}

}  // namespace rcot
