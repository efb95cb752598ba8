// common.cuh -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mucon_b200.h"

namespace mucon {

// Records the last CUDA error string for mucon_last_cuda_error().
void set_cuda_error(cudaError_t e, const char* where);

#define MUCON_CUDA_CHECK(expr)                                   \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) {                                     \
      ::mucon::set_cuda_error(_e, #expr);                        \
      return MUCON_ECUDA;                                        \
    }                                                            \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier (shared::cta) -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// the same on a 32-bit shared-space address (hot loops keep these instead of generic pointers: a
// generic pointer is converted -- S2UR + ULEA + LEA -- at every use)
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ float ld_shared(uint32_t addr, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double ld_shared(uint32_t addr, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP) -------------
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- Ampere-style cp.async for small gathers --------------------------------------------------
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// -0 is the exact additive identity of IEEE addition (x + -0 == x for every x, including -0)
template <typename T>
__device__ __forceinline__ T neg_zero();
template <>
__device__ __forceinline__ float neg_zero<float>() { return -0.0f; }
template <>
__device__ __forceinline__ double neg_zero<double>() { return -0.0; }

// Activation modes of the backbone kernels' relu flags: 0 none, 1 ReLU, 2 leaky ReLU with torch's default slope 0.01
// (model.ft.leaky_relu, reference src/core/modules/temporal.py:35-41,98-101).
// Branch-free on purpose: written as nested conditionals this compiled to two BRANCHES per element inside the GEMM
// epilogues (conv_gemm_kernel went from 2.9 to 9.0 us per 128-row tile).  max(v, 0) + slope * min(v, 0) is exact for
// both activations (slope 0: ReLU, the second term is +-0; slope 0.01: leaky ReLU, the first term is 0 for v < 0).
__device__ __forceinline__ float act_mode(float v, int mode) {
  const float slope = mode == 2 ? 0.01f : 0.f;
  const float a = fmaxf(v, 0.f) + slope * fminf(v, 0.f);
  return mode == 0 ? v : a;
}

// Result stores of the alignment kernels: the local payload buffer and, when a multi-GPU result exchange is set up
// (mucon_viterbi_batch.peer_delta), the same location of this rank's slot in every peer's receive buffer.
__device__ __forceinline__ void put_score(const mucon_viterbi_batch& b, int64_t u, double v) {
  b.score[u] = v;
  for (int p = 0; p < b.n_peers; ++p)
    *reinterpret_cast<double*>(reinterpret_cast<char*>(b.score + u) + b.peer_delta[p]) = v;
}
__device__ __forceinline__ void put_seg(const mucon_viterbi_batch& b, int64_t i, int32_t v) {
  b.seg_blocks[i] = v;
  for (int p = 0; p < b.n_peers; ++p)
    *reinterpret_cast<int32_t*>(reinterpret_cast<char*>(b.seg_blocks + i) + b.peer_delta[p]) = v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace mucon
