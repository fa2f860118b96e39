// Error reporting and device checks of the C ABI (include/mpsnerf.h).
#include <stdarg.h>

#include "common.cuh"

namespace mps {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mps

extern "C" const char* mpsnerf_last_error(void) { return mps::g_err; }

extern "C" int mpsnerf_abi_version(void) { return MPSNERF_ABI_VERSION; }

extern "C" int mpsnerf_check_device(int dev) {
  cudaDeviceProp p;
  MPS_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    mps::set_error("device %d is sm_%d%d; libmpsnerf_b200 is built for sm_100a only", dev, p.major, p.minor);
    return MPSNERF_EARCH;
  }
  return MPSNERF_OK;
}
