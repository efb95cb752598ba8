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
    case MUCON_ESHAPE: return "input larger than the session was created for";
    default: return "unknown error";
  }
}

extern "C" const char* mucon_last_cuda_error(void) { return mucon::g_err; }

// Optional cap on the grid of the persistent kernels (thread-local: set by the launching host thread, read by the
// launchers through mucon_device_sm_count on the same thread).  Lets two kernels that each want one CTA per SM run
// SIDE BY SIDE on disjoint SM sets from two streams -- the HBM-bound projection of video chunk k+1 next to the
// tensor-bound layer kernels of chunk k (MuConBackbone.infer_pooled_pipelined).
static thread_local int t_sm_limit = 0;
extern "C" int mucon_set_sm_limit(int n) {
  const int old = t_sm_limit;
  t_sm_limit = n > 0 ? n : 0;
  return old;
}

extern "C" int mucon_device_sm_count(void) {
  // cached per device (read-only after the first query; a benign race writes the same value twice)
  static int cache[64] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64 && cache[dev]) n = cache[dev];
  else {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64) cache[dev] = n;
  }
  return (t_sm_limit > 0 && t_sm_limit < n) ? t_sm_limit : n;
}

// ---- receive buffers of the multi-GPU result exchange (CUDA IPC; see mucon_viterbi_batch.peer_delta) ------------
extern "C" int mucon_peer_alloc(size_t bytes, void** dev_ptr_out, unsigned char handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!dev_ptr_out || !handle_out || bytes == 0) return MUCON_EINVAL;
  void* p = nullptr;
  MUCON_CUDA_CHECK(cudaMalloc(&p, bytes));
  MUCON_CUDA_CHECK(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    mucon::set_cuda_error(e, "cudaIpcGetMemHandle");
    return MUCON_ECUDA;
  }
  memcpy(handle_out, &h, 64);
  *dev_ptr_out = p;
  return MUCON_OK;
}

extern "C" int mucon_peer_open(const unsigned char handle[64], void** dev_ptr_out) {
  if (!handle || !dev_ptr_out) return MUCON_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  MUCON_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr_out = p;
  return MUCON_OK;
}

extern "C" int mucon_peer_close(void* dev_ptr) {
  if (!dev_ptr) return MUCON_EINVAL;
  MUCON_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
  return MUCON_OK;
}

extern "C" int mucon_peer_free(void* dev_ptr) {
  if (!dev_ptr) return MUCON_EINVAL;
  MUCON_CUDA_CHECK(cudaFree(dev_ptr));
  return MUCON_OK;
}
