// api.cu -- ABI version, error strings, device queries.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace mucon {
static thread_local char g_err[512] = "";
void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
}
}  // namespace mucon

extern "C" int mucon_abi_version(void) { return MUCON_ABI_VERSION; }

extern "C" const char* mucon_strerror(int code) {
  switch (code) {
    case MUCON_OK: return "ok";
    case MUCON_EINVAL: return "invalid argument";
    case MUCON_EUNSUPPORTED: return "shape not supported by the sm_100a kernels";
    case MUCON_ECUDA: return "CUDA runtime error";
    case MUCON_EALIGN: return "misaligned pointer";
    default: return "unknown error";
  }
}

extern "C" const char* mucon_last_cuda_error(void) { return mucon::g_err; }

extern "C" int mucon_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}
