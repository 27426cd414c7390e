// api.cu -- library-wide C-ABI plumbing: version, last-error string, device check.
#include <stdarg.h>

#include "common.cuh"

namespace rcot {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return RCOT_ERR_CUDA;
  }
  return RCOT_OK;
}

}  // namespace rcot

extern "C" {

int rcot_version(void) { return 100; }

const char* rcot_last_error(void) { return rcot::g_err; }

// 0 when the current device is sm_100 (B200); RCOT_ERR_ARCH otherwise. There is no fallback path.
int rcot_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    rcot::set_error("rcot_check_device: no CUDA device");
    cudaGetLastError();
    return RCOT_ERR_CUDA;
  }
  if (prop.major != 10) {
    rcot::set_error("rcot_check_device: device is sm_%d%d, this library is sm_100a only", prop.major, prop.minor);
    return RCOT_ERR_ARCH;
  }
  return RCOT_OK;
}

int rcot_zero(void* ptr, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return RCOT_OK;
  if (ptr == nullptr) {
    rcot::set_error("rcot_zero: null pointer");
    return RCOT_ERR_ARG;
  }
  cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, stream);
  if (e != cudaSuccess) {
    rcot::set_error("rcot_zero: %s", cudaGetErrorString(e));
    return RCOT_ERR_CUDA;
  }
  return RCOT_OK;
}

}  // extern "C"
