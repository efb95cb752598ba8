// backbone_gemm.cuh -- tcgen05 (5th-gen tensor core) GEMM for the backbone's dense contractions.
//
//   out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] )        A: [M, K] fp32, W: [N = 128, K] fp32
//
// used for the 2048 -> 128 input projection of reference src/core/modules/temporal.py:133
// (`first_conv`, a 1x1 Conv1d == a GEMM over frames).  The fp32 operands are consumed as TF32 by
// `tcgen05.mma.kind::tf32` straight from shared memory: no conversion pass over the 8 KB/frame
// feature stream, which is what bounds this op (HBM).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: 2-D tiled `cp.async.bulk.tensor` loads of a [128 x 32] fp32 A tile and
//               a [128 x 32] W tile (128-byte rows, SWIZZLE_128B) per stage, completion on mbarriers
//   warp 1      allocates TMEM (2 x 128 columns: double-buffered 128x128 fp32 accumulators) and
//               issues the MMAs (one elected lane): 4 x (M128 N128 K8) per stage, `tcgen05.commit`
//               releases the stage / publishes the accumulator
//   warps 2-5   epilogue: `tcgen05.ld` 32 lanes x 32 columns, + bias, ReLU, 16-byte stores
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mucon {
namespace gemm {

constexpr int BM = 128, BN = 128, BK = 32;    // BK fp32 = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;                     // tf32: 32 bytes of K per instruction
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256;   // both proj_gemm_kernel variants: 6 x 32 KB = 4 x 48 KB

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp layout):
// start address >> 4 in [0,14), LBO >> 4 in [16,30) (unused for swizzled K-major), SBO >> 4 in
// [32,46) = 1024 B between 8-row groups, version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
// N >> 3 in [17,23), M >> 4 in [24,29)
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// The same load delivered to every CTA of the cluster named in cta_mask, at the same shared-memory
// offset, signalling the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// tcgen05.commit arriving on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Conv1d zero padding inside an operand tile whose rows are 128-byte lines (the swizzle only permutes the 16-byte
// chunks INSIDE a line): rows [0, lo) and [hi, rows) of the tile lie outside the video and are cleared.  Consecutive
// lanes take consecutive 16-byte chunks (a warp store covers four whole rows, 512 contiguous bytes: no bank
// conflicts) -- a lane per row would put all 32 lanes on the same four banks, 32-way serialised, which made the
// fix-up warp the longest link of the per-tile chain wherever most tiles are partly filled (the T/8 and T/16 levels).
__device__ __forceinline__ void zero_pad_rows(unsigned char* tile, int lo, int hi, int rows, int lane) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  const int e0 = (lo < rows ? (lo > 0 ? lo : 0) : rows) * 128;
  for (int o = lane * 16; o < e0; o += 512) *reinterpret_cast<uint4*>(tile + o) = z;
  const int b1 = (hi > 0 ? hi : 0) * 128, e1 = rows * 128;
  for (int o = b1 + lane * 16; o < e1; o += 512) *reinterpret_cast<uint4*>(tile + o) = z;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MT = 128-row M tiles per CTA tile.  With MT = 2 a weight k-block staged in shared memory is used for 256 rows:
// the kernel's traffic from L2 into the SMs drops from 2x to 1.5x the feature bytes.  That matters because the
// MT = 1 kernel ran AT the L2's total delivery rate (63 GB in 5.3 ms = 11.9 TB/s = 6300 B/cycle at 1.9 GHz, the
// measured LTS cap), half of it weight tiles that never change, while HBM itself was at 5.9 TB/s.
template <int MT>
__global__ void __launch_bounds__(THREADS, 1)
proj_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const float* __restrict__ bias, float* __restrict__ out, int M, int K, int relu, int out_bf16) {
  constexpr int PSTAGES = MT == 1 ? STAGES : 4;
  constexpr int PA_BYTES = MT * A_BYTES, PSTAGE_BYTES = PA_BYTES + B_BYTES;
  constexpr int ACC_COLS = MT * BN;            // TMEM columns of one accumulator set
  constexpr int TILE_M = MT * BM;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned stage buffers (SWIZZLE_128B atoms are 8 rows x 128 B)
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_mem = base;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + PSTAGES * PSTAGE_BYTES);
  uint64_t* empty = full + PSTAGES;
  uint64_t* tfull = empty + PSTAGES;   // [2] accumulator ready
  uint64_t* tempty = tfull + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (M + TILE_M - 1) / TILE_M;
  const int num_kb = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_fence_init();
  }
  if (warp == 1) {  // TMEM: two sets of MT 128x128 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * ACC_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = tile * TILE_M;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], PSTAGE_BYTES);
          unsigned char* a = stage_mem + s * PSTAGE_BYTES;
          tma_load_2d(a, &tmA, kb * BK, m0, &full[s]);          // box = TILE_M rows: MT consecutive [128 x 32] tiles
          tma_load_2d(a + PA_BYTES, &tmB, kb * BK, 0, &full[s]);
          if (++s == PSTAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = instr_desc_tf32(BM, BN);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_ph ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(stage_mem + s * PSTAGE_BYTES);
          const uint64_t bdesc = smem_desc(a_addr + PA_BYTES);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = smem_desc(a_addr + mt * A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 32 bytes of K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
              mma_tf32(d_tmem + mt * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
          }
          mma_commit(&empty[s]);                       // frees the stage when these MMAs retire
          if (kb == num_kb - 1) mma_commit(&tfull[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++s == PSTAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
      const int row = tile * TILE_M + mt * BM + q * 32 + lane;
      float* orow = out + static_cast<int64_t>(row) * BN;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ACC_COLS + mt * BN + c * 32, r);
        if (row < M && out_bf16) {
          // 16-bit activations (1: bf16, 2: fp16) for the layer kernels of backbone_bf16.cuh: 32 columns = 64 bytes = two
          // 32-byte stores
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float lo = __uint_as_float(r[2 * j + 0]) + __ldg(bias + c * 32 + 2 * j + 0);
            float hi = __uint_as_float(r[2 * j + 1]) + __ldg(bias + c * 32 + 2 * j + 1);
            lo = act_mode(lo, relu); hi = act_mode(hi, relu);
            if (out_bf16 == 2) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(hi), "f"(lo));
            else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(hi), "f"(lo));
          }
          unsigned short* ob = reinterpret_cast<unsigned short*>(out) + static_cast<int64_t>(row) * BN + c * 32;
#pragma unroll
          for (int g = 0; g < 2; ++g)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"l"(ob + 16 * g), "r"(pk[8 * g + 0]), "r"(pk[8 * g + 1]), "r"(pk[8 * g + 2]), "r"(pk[8 * g + 3]),
                           "r"(pk[8 * g + 4]), "r"(pk[8 * g + 5]), "r"(pk[8 * g + 6]), "r"(pk[8 * g + 7])
                         : "memory");
        } else if (row < M) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v;
            v.x = __uint_as_float(r[j + 0]) + __ldg(bias + c * 32 + j + 0);
            v.y = __uint_as_float(r[j + 1]) + __ldg(bias + c * 32 + j + 1);
            v.z = __uint_as_float(r[j + 2]) + __ldg(bias + c * 32 + j + 2);
            v.w = __uint_as_float(r[j + 3]) + __ldg(bias + c * 32 + j + 3);
            v.x = act_mode(v.x, relu); v.y = act_mode(v.y, relu); v.z = act_mode(v.z, relu); v.w = act_mode(v.w, relu);
            *reinterpret_cast<float4*>(orow + c * 32 + j) = v;
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * ACC_COLS));
  }
}

}  // namespace gemm

// =============================================================================================
// Temporal convolution as a tcgen05 GEMM (k = 1 or k = 3 dilated, 128 -> 128 channels).
//
//   out[t, :] = act( sum_tap X[t + (tap - taps/2)*dil, :] . W[tap]^T + bias ) (+ residual) (ReLU)
//
// For a fixed tap the rows of a [128 time steps x 128 channels] activation tile shifted by
// (tap-1)*dil are again consecutive rows of the time-major activation tensor, so each tap is four
// more k-blocks of the same accumulator: A tiles are TMA loads at a shifted row coordinate, B tiles
// are the tap's [Cout x Cin] weight slice.  Zero padding at a video's ends (Conv1d(padding=dil),
// reference src/core/modules/temporal.py:21-27): taps that only see padding are skipped, rows of a
// tile that fall outside the video are zeroed in shared memory by a fix-up warp that sits between
// the TMA producer and the MMA issuer (generic-proxy writes are made visible to the tensor core's
// async proxy with fence.proxy.async before the stage is handed on).
// 224 threads: warp 0 producer, warp 1 MMA + TMEM, warp 2 fix-up, warps 3-6 epilogue.
namespace convgemm {

using namespace gemm;
constexpr int CTHREADS = 224;
constexpr int C = 128;          // channels in = channels out
constexpr int KB_PER_TAP = C / BK;
constexpr int CSTAGES = 4;
constexpr int EPI_LD = C + 4;   // padded row of the epilogue staging tile (floats): 16-byte aligned, and a
                                // warp-wide 16-byte store to 32 rows spreads over all banks
constexpr int EPI_BYTES = 32 * EPI_LD * 4;  // per epilogue warp
constexpr int CSMEM_BYTES = 1024 + CSTAGES * STAGE_BYTES + 4 * EPI_BYTES + 512;

struct Tile {
  long long row0;  // global row of the video's first time step at this resolution
  int t0;          // first time step of the tile within the video
  int T;           // video length at this resolution
};

__device__ __forceinline__ bool tap_live(int shift, int T) { return shift < T && -shift < T; }

// Row shifts of the taps (time steps relative to the output row); one of them must be 0.  A k = 3
// dilated Conv1d is {-d, 0, +d}; two dilated convolutions folded through a 1x1 fusion conv (MS-TCN++
// first stage, temporal.py:150-204) are {-d1, -d2, 0, +d2, +d1}.
constexpr int kMaxTaps = 6;
struct TapShifts {
  int n;
  int s[kMaxTaps];
};

// Epilogue phase A of conv_gemm_kernel: TMEM (lane = row) -> + bias -> activation -> padded shared tile.  The bias sits
// in shared memory (zeros when the launch has none) and is read four values at a time.
template <int MODE>
__device__ __forceinline__ void epi_phase_a(uint32_t taddr, const float* bias_s, float* et, int lane) {
#pragma unroll
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(taddr + c * 32, r);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
      float4 v = make_float4(__uint_as_float(r[j]) + b4.x, __uint_as_float(r[j + 1]) + b4.y,
                             __uint_as_float(r[j + 2]) + b4.z, __uint_as_float(r[j + 3]) + b4.w);
      if (MODE != 0) { v.x = act_mode(v.x, MODE); v.y = act_mode(v.y, MODE); v.z = act_mode(v.z, MODE); v.w = act_mode(v.w, MODE); }
      *reinterpret_cast<float4*>(et + lane * EPI_LD + c * 32 + j) = v;
    }
  }
}

__global__ void __launch_bounds__(CTHREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const Tile* __restrict__ tiles, int num_tiles, const TapShifts ts, const float* __restrict__ bias,
                 const float* __restrict__ residual, float* __restrict__ out, int relu_mid, int relu_final,
                 const float* __restrict__ mul = nullptr, const float* __restrict__ gate = nullptr) {
  // Epilogue order: v = acc (+ bias) -> ReLU (relu_mid) -> * mul (dropout mask of the forward pass) -> + residual ->
  // ReLU (relu_final) -> gate (v = gate > 0 ? v : 0: the ReLU derivative of the backward pass).  bias / residual /
  // mul / gate may be null.
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_mem = base;
  float* epi_mem = reinterpret_cast<float*>(base + CSTAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + CSTAGES * STAGE_BYTES + 4 * EPI_BYTES);
  uint64_t* ready = full + CSTAGES;
  uint64_t* empty = ready + CSTAGES;
  uint64_t* tfull = empty + CSTAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  __shared__ __align__(16) float bias_s[C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < C) bias_s[threadIdx.x] = bias ? __ldg(bias + threadIdx.x) : 0.f;   // published by the __syncthreads below
  if (threadIdx.x == 0) {
    for (int s = 0; s < CSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int taps = ts.n;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int ti = blockIdx.x; ti < num_tiles; ti += gridDim.x) {
        const Tile tl = tiles[ti];
        for (int tap = 0; tap < taps; ++tap) {
          const int shift = ts.s[tap];
          if (!tap_live(shift, tl.T)) continue;
          const int row = static_cast<int>(tl.row0) + tl.t0 + shift;  // may be negative: TMA zero-fills
          for (int kc = 0; kc < KB_PER_TAP; ++kc) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
            unsigned char* a = stage_mem + s * STAGE_BYTES;
            tma_load_2d(a, &tmX, kc * BK, row, &full[s]);
            tma_load_2d(a + A_BYTES, &tmW, kc * BK, tap * C, &full[s]);
            if (++s == CSTAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ fix-up warp =================================
    int s = 0;
    uint32_t ph = 0;
    for (int ti = blockIdx.x; ti < num_tiles; ti += gridDim.x) {
      const Tile tl = tiles[ti];
      for (int tap = 0; tap < taps; ++tap) {
        const int shift = ts.s[tap];
        if (!tap_live(shift, tl.T)) continue;
        const int lo = -(tl.t0 + shift);          // rows r < lo are before the video
        const int hi = tl.T - (tl.t0 + shift);    // rows r >= hi are after it
        const bool fix = lo > 0 || hi < BM;
        for (int kc = 0; kc < KB_PER_TAP; ++kc) {
          mbar_wait(&full[s], ph);
          if (fix) {
            zero_pad_rows(stage_mem + s * STAGE_BYTES, lo, hi, BM, lane);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[s]);
          if (++s == CSTAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = instr_desc_tf32(BM, BN);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int ti = blockIdx.x; ti < num_tiles; ti += gridDim.x, ++it) {
      const Tile tl = tiles[ti];
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      int issued = 0;
      for (int tap = 0; tap < taps; ++tap) {
        const int shift = ts.s[tap];
        if (!tap_live(shift, tl.T)) continue;
        for (int kc = 0; kc < KB_PER_TAP; ++kc) {
          mbar_wait(&ready[s], ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_u32(stage_mem + s * STAGE_BYTES);
            const uint64_t adesc = smem_desc(a_addr), bdesc = smem_desc(a_addr + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              mma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (issued | k) != 0);
            mma_commit(&empty[s]);
          }
          __syncwarp();
          ++issued;
          if (++s == CSTAGES) { s = 0; ph ^= 1; }
        }
      }
      if (lane == 0) mma_commit(&tfull[acc]);  // the centre tap is always live, so issued > 0
      __syncwarp();
    }
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;
    int it = 0;
    for (int ti = blockIdx.x; ti < num_tiles; ti += gridDim.x, ++it) {
      const Tile tl = tiles[ti];
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      // phase A: TMEM -> registers (lane = row) -> + bias (ReLU) -> padded shared tile
      float* et = epi_mem + (warp - 3) * (EPI_BYTES / 4);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      // one uniform branch per tile, the activation compiled into each variant (not a test per element)
      if (relu_mid == 0) epi_phase_a<0>(taddr, bias_s, et, lane);
      else if (relu_mid == 1) epi_phase_a<1>(taddr, bias_s, et, lane);
      else epi_phase_a<2>(taddr, bias_s, et, lane);
      // the accumulator is drained: hand it back before the (slower) global phase
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      // phase B: one row per iteration, 32 lanes x 16 bytes = a full 512-byte row: coalesced
      // residual read and store
      const int tbase = tl.t0 + q * 32;
      const long long rbase = tl.row0 + tbase;
      const int nrow = min(32, tl.T - tbase);  // rows of this warp inside the video (may be <= 0)
#pragma unroll
      for (int r0 = 0; r0 < 32; r0 += 8) {
        float4 rv[8], v[8];
        if (residual) {  // eight independent 512-byte row loads in flight
#pragma unroll
          for (int e = 0; e < 8; ++e)
            rv[e] = (r0 + e < nrow) ? __ldg(reinterpret_cast<const float4*>(residual + (rbase + r0 + e) * C) + lane)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = *reinterpret_cast<const float4*>(et + (r0 + e) * EPI_LD + lane * 4);
        if (mul) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (r0 + e < nrow) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(mul + (rbase + r0 + e) * C) + lane);
              v[e].x *= m.x; v[e].y *= m.y; v[e].z *= m.z; v[e].w *= m.w;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (residual) { v[e].x += rv[e].x; v[e].y += rv[e].y; v[e].z += rv[e].z; v[e].w += rv[e].w; }
          if (relu_final) {
            v[e].x = act_mode(v[e].x, relu_final); v[e].y = act_mode(v[e].y, relu_final);
            v[e].z = act_mode(v[e].z, relu_final); v[e].w = act_mode(v[e].w, relu_final);
          }
        }
        if (gate) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (r0 + e < nrow) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (rbase + r0 + e) * C) + lane);
              v[e].x = g.x > 0.f ? v[e].x : 0.f; v[e].y = g.y > 0.f ? v[e].y : 0.f;
              v[e].z = g.z > 0.f ? v[e].z : 0.f; v[e].w = g.w > 0.f ? v[e].w : 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (r0 + e < nrow) *(reinterpret_cast<float4*>(out + (rbase + r0 + e) * C) + lane) = v[e];
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

}  // namespace convgemm

// =============================================================================================
// One whole WaveNet layer per launch (reference src/core/modules/temporal.py:43-53 + the optional
// max_pool1d(2) of :137-139):   y = relu(conv_k3_dil(x) + bd);  out = conv_1x1(y) + b1 + x;  [pool]
//
// Per 128-row tile: GEMM 1 (3 taps x 4 k-blocks, as conv_gemm_kernel) accumulates into TMEM
// columns 0-127; the epilogue warps turn that accumulator into relu(. + bd) and write it to shared
// memory directly in the K-major SWIZZLE_128B layout the tensor core reads (four [128 x 32] sub-tiles),
// so it becomes the A operand of GEMM 2 (4 k-blocks against the 1x1 weights, TMEM columns 128-255)
// without ever leaving the SM; the second epilogue adds bias and the residual, optionally max-pools
// adjacent rows, and writes coalesced 512-byte rows.  The intermediate activation (and the separate
// pooling pass) cost no HBM traffic.
// 384 threads: warp 0 producer, warp 1 MMA + TMEM, warp 2 fix-up, warp 3 idle, warps 4-11 epilogue.
namespace layer {

// Optional per-role timeline of CTA 0 (developer builds with -DMUCON_LAYER_TRACE, see scripts/trace_layer.py):
// clock64() stamps per tile for the MMA warp and epilogue warp 4 of wavenet_layer_kernel.  This is how the
// epilogue warps were found to be the bottleneck of the layer kernel.  Compiles to nothing by default.
#ifdef MUCON_LAYER_TRACE
__device__ long long g_trace[32][128];
#define MUCON_TR(ev, it) do { if (blockIdx.x == 0 && (it) < 128) g_trace[ev][it] = clock64(); } while (0)
#else
#define MUCON_TR(ev, it) do { } while (0)
#endif


using namespace gemm;
constexpr int LTHREADS = 384;  // producer, MMA, fix-up, (idle), 8 epilogue warps
constexpr int EPI_WARPS = 8;   // two warps per TMEM lane quarter, 64 accumulator columns each
constexpr int C = 128;
constexpr int KB_PER_TAP = C / BK;
constexpr int LSTAGES = 4;
constexpr int Y_BYTES = BM * C * 4;  // 64 KB: A operand of GEMM 2, later the epilogue staging area
constexpr int LSMEM_BYTES = 1024 + LSTAGES * STAGE_BYTES + Y_BYTES + 512;

struct Tile {
  long long row0;      // first row of the video at the input resolution
  long long row0_out;  // first row of the video in the output tensor (differs when pooling)
  int t0;              // first time step of the tile within the video
  int T;               // video length at the input resolution
};

__device__ __forceinline__ bool tap_live(int shift, int T) { return shift < T && -shift < T; }

// PAIR: two CTAs of a cluster work on adjacent tiles of the same video in lockstep and share the
// weight traffic: each loads half of every weight k-block (64 of the 128 output-channel rows) and
// multicasts it to both (tmWd / tmW1 are then maps with 64-row boxes), so a tile costs 128 KB of
// weight reads from L2 instead of 256 KB next to its 64 KB of activations.  A stage may be refilled
// only when BOTH CTAs' MMAs have released it: the stage-release commit is multicast to both CTAs
// and the `empty` barriers count two arrivals.  The tile list holds the two tiles of a pair next to
// each other (the host pads every video to an even number of tiles; a padding tile starts at t0 >= T,
// so all of its rows are zeroed by the fix-up warp and none is stored).
template <bool PAIR>
__global__ void __launch_bounds__(LTHREADS, 1)
wavenet_layer_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWd,
                     const __grid_constant__ CUtensorMap tmW1, const Tile* __restrict__ tiles, int num_tiles, int dil,
                     const float* __restrict__ bd, const float* __restrict__ b1, const float* __restrict__ x,
                     float* __restrict__ out, int pool, int relu_final) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_mem = base;
  unsigned char* ybuf = base + LSTAGES * STAGE_BYTES;  // 1024-byte aligned (STAGE_BYTES is a multiple of 1024)
  uint64_t* full = reinterpret_cast<uint64_t*>(ybuf + Y_BYTES);
  uint64_t* ready = full + LSTAGES;
  uint64_t* empty = ready + LSTAGES;
  uint64_t* a1full = empty + LSTAGES;   // accumulator 1 complete
  uint64_t* a1empty = a1full + 1;       // accumulator 1 drained by the epilogue
  uint64_t* yready = a1empty + 1;       // Y written and published to the async proxy
  uint64_t* a2full = yready + 1;        // accumulator 2 complete (Y no longer read)
  uint64_t* a2empty = a2full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < LSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 1); mbar_init(&empty[s], PAIR ? 2 : 1); }
    mbar_init(a1full, 1); mbar_init(a1empty, EPI_WARPS); mbar_init(yready, EPI_WARPS); mbar_init(a2full, 1); mbar_init(a2empty, EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tile walk: alone, CTA b takes tiles b, b + grid, ...; paired, cluster c takes tile pairs c, c + clusters, ...
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile_first = PAIR ? 2 * static_cast<int>(blockIdx.x >> 1) + rank : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(gridDim.x & ~1u) : static_cast<int>(gridDim.x);
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWd) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int ti = tile_first; ti < num_tiles; ti += tile_step) {
        const Tile tl = tiles[ti];
        for (int tap = 0; tap < 3; ++tap) {
          const int shift = (tap - 1) * dil;
          if (!tap_live(shift, tl.T)) continue;
          const int row = static_cast<int>(tl.row0) + tl.t0 + shift;
          for (int kc = 0; kc < KB_PER_TAP; ++kc) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
            unsigned char* a = stage_mem + s * STAGE_BYTES;
            tma_load_2d(a, &tmX, kc * BK, row, &full[s]);
            if (PAIR) tma_load_2d_mc(a + A_BYTES + rank * (B_BYTES / 2), &tmWd, kc * BK, tap * C + rank * (BN / 2), &full[s], 3);
            else tma_load_2d(a + A_BYTES, &tmWd, kc * BK, tap * C, &full[s]);
            if (++s == LSTAGES) { s = 0; ph ^= 1; }
          }
        }
        for (int kc = 0; kc < KB_PER_TAP; ++kc) {  // 1x1 weights: B half of the stage only
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], B_BYTES);
          unsigned char* bdst = stage_mem + s * STAGE_BYTES + A_BYTES;
          if (PAIR) tma_load_2d_mc(bdst + rank * (B_BYTES / 2), &tmW1, kc * BK, rank * (BN / 2), &full[s], 3);
          else tma_load_2d(bdst, &tmW1, kc * BK, 0, &full[s]);
          if (++s == LSTAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ fix-up warp =================================
    int s = 0;
    uint32_t ph = 0;
    for (int ti = tile_first; ti < num_tiles; ti += tile_step) {
      const Tile tl = tiles[ti];
      for (int tap = 0; tap < 3; ++tap) {
        const int shift = (tap - 1) * dil;
        if (!tap_live(shift, tl.T)) continue;
        const int lo = -(tl.t0 + shift);
        const int hi = tl.T - (tl.t0 + shift);
        const bool fix = lo > 0 || hi < BM;
        for (int kc = 0; kc < KB_PER_TAP; ++kc) {
          mbar_wait(&full[s], ph);
          if (fix) {
            zero_pad_rows(stage_mem + s * STAGE_BYTES, lo, hi, BM, lane);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[s]);
          if (++s == LSTAGES) { s = 0; ph ^= 1; }
        }
      }
      for (int kc = 0; kc < KB_PER_TAP; ++kc) {
        mbar_wait(&full[s], ph);
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
        if (++s == LSTAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = instr_desc_tf32(BM, BN);
    int s = 0, it = 0;
    uint32_t ph = 0;
    const uint32_t d1 = tmem_base, d2 = tmem_base + BN;
    for (int ti = tile_first; ti < num_tiles; ti += tile_step, ++it) {
      const Tile tl = tiles[ti];
      const uint32_t tph = it & 1;
      // ---- GEMM 1: dilated conv
      mbar_wait(a1empty, tph ^ 1);
      tc_fence_after();
      if (lane == 0) MUCON_TR(2, it);  // GEMM 1 may start
      int issued = 0;
      for (int tap = 0; tap < 3; ++tap) {
        const int shift = (tap - 1) * dil;
        if (!tap_live(shift, tl.T)) continue;
        for (int kc = 0; kc < KB_PER_TAP; ++kc) {
          mbar_wait(&ready[s], ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_u32(stage_mem + s * STAGE_BYTES);
            const uint64_t adesc = smem_desc(a_addr), bdesc = smem_desc(a_addr + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) mma_tf32(d1, adesc + 2 * k, bdesc + 2 * k, idesc, (issued | k) != 0);
            if (PAIR) mma_commit_mc(&empty[s], 3); else mma_commit(&empty[s]);
          }
          __syncwarp();
          ++issued;
          if (++s == LSTAGES) { s = 0; ph ^= 1; }
        }
      }
      if (lane == 0) { mma_commit(a1full); MUCON_TR(3, it); }  // GEMM 1 issued
      __syncwarp();
      // ---- GEMM 2: 1x1 conv on relu(acc1 + bd), which the epilogue warps wrote to ybuf
      mbar_wait(yready, tph);
      mbar_wait(a2empty, tph ^ 1);
      tc_fence_after();
      if (lane == 0) MUCON_TR(4, it);  // GEMM 2 may start
      for (int kc = 0; kc < KB_PER_TAP; ++kc) {
        mbar_wait(&ready[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = smem_desc(smem_u32(ybuf + kc * A_BYTES));
          const uint64_t bdesc = smem_desc(smem_u32(stage_mem + s * STAGE_BYTES + A_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) mma_tf32(d2, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
          if (PAIR) mma_commit_mc(&empty[s], 3); else mma_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == LSTAGES) { s = 0; ph ^= 1; }
      }
      if (lane == 0) { mma_commit(a2full); MUCON_TR(5, it); }  // GEMM 2 issued
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    // Eight epilogue warps: the per-tile timeline (clock stamps, scratch build) showed the four epilogue
    // warps busy ~12.5 k of the 14.4 k cycles of a tile (8.5 k of it in the residual-load / store phase)
    // while the MMA and TMA warps waited for them.  Two warps now share a TMEM lane quarter: each takes
    // 64 of the 128 accumulator columns in the TMEM phases and 16 of the quarter's 32 rows in the
    // coalesced phase.
    const int q = warp & 3;            // TMEM lane quarter (a warp may only touch lanes 32*(warp%4) ..)
    const int half = (warp - 4) >> 2;  // which 64 accumulator columns / which 16 rows of the quarter
    const int r = q * 32 + lane;       // tile row owned in the TMEM-load phases
    int it = 0;
    for (int ti = tile_first; ti < num_tiles; ti += tile_step, ++it) {
      const Tile tl = tiles[ti];
      const uint32_t tph = it & 1;
      // ---- epilogue 1: acc1 -> relu(. + bd) -> ybuf in the K-major SWIZZLE_128B operand layout
      mbar_wait(a1full, tph);
      tc_fence_after();
      if (warp == 4 && lane == 0) MUCON_TR(6, it);  // epilogue 1 starts
#pragma unroll
      for (int kh = 0; kh < BN / 64; ++kh) {
        const int kc = half * (BN / 64) + kh;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kc * 32, v);
        unsigned char* rowp = ybuf + kc * A_BYTES + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = fmaxf(__uint_as_float(v[4 * j + 0]) + __ldg(bd + kc * 32 + 4 * j + 0), 0.f);
          o.y = fmaxf(__uint_as_float(v[4 * j + 1]) + __ldg(bd + kc * 32 + 4 * j + 1), 0.f);
          o.z = fmaxf(__uint_as_float(v[4 * j + 2]) + __ldg(bd + kc * 32 + 4 * j + 2), 0.f);
          o.w = fmaxf(__uint_as_float(v[4 * j + 3]) + __ldg(bd + kc * 32 + 4 * j + 3), 0.f);
          *reinterpret_cast<float4*>(rowp + ((j ^ (r & 7)) << 4)) = o;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(a1empty); mbar_arrive(yready); }
      if (warp == 4 && lane == 0) MUCON_TR(7, it);  // epilogue 1 done
      // ---- epilogue 2: acc2 + b1 -> staging (the ybuf bytes, free once GEMM 2 has completed)
      mbar_wait(a2full, tph);
      tc_fence_after();
      if (warp == 4 && lane == 0) MUCON_TR(8, it);  // epilogue 2 starts
      float* et = reinterpret_cast<float*>(ybuf) + q * (32 * C);  // [32 rows][128], 16-byte chunks XOR-swizzled by row
#pragma unroll
      for (int kh = 0; kh < BN / 64; ++kh) {
        const int kc = half * (BN / 64) + kh;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + BN + kc * 32, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = __uint_as_float(v[4 * j + 0]) + __ldg(b1 + kc * 32 + 4 * j + 0);
          o.y = __uint_as_float(v[4 * j + 1]) + __ldg(b1 + kc * 32 + 4 * j + 1);
          o.z = __uint_as_float(v[4 * j + 2]) + __ldg(b1 + kc * 32 + 4 * j + 2);
          o.w = __uint_as_float(v[4 * j + 3]) + __ldg(b1 + kc * 32 + 4 * j + 3);
          const int chunk = kc * 8 + j;
          *reinterpret_cast<float4*>(et + lane * C + ((chunk ^ (lane & 7)) << 2)) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2empty);
      // both warps of a quarter have staged their 64 columns before either reads whole rows
      named_bar_sync(3 + q, 64);
      if (warp == 4 && lane == 0) MUCON_TR(9, it);  // accumulator 2 staged
      // ---- coalesced output: + residual, optional ReLU, optional max-pool of adjacent rows;
      // the quarter's 32 rows are split between its two warps
      const int tbase = tl.t0 + q * 32;
      const int nrow = min(32, tl.T - tbase);
      const long long rbase = tl.row0 + tbase;
      if (!pool) {
#pragma unroll
        for (int r0 = half * 16; r0 < half * 16 + 16; r0 += 8) {
          float4 rv[8], v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            rv[e] = (r0 + e < nrow) ? __ldg(reinterpret_cast<const float4*>(x + (rbase + r0 + e) * C) + lane)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            v[e] = *reinterpret_cast<const float4*>(et + (r0 + e) * C + ((lane ^ ((r0 + e) & 7)) << 2));
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[e].x += rv[e].x; v[e].y += rv[e].y; v[e].z += rv[e].z; v[e].w += rv[e].w;
            if (relu_final) {
              v[e].x = fmaxf(v[e].x, 0.f); v[e].y = fmaxf(v[e].y, 0.f); v[e].z = fmaxf(v[e].z, 0.f); v[e].w = fmaxf(v[e].w, 0.f);
            }
            if (r0 + e < nrow) *(reinterpret_cast<float4*>(out + (tl.row0_out + tbase + r0 + e) * C) + lane) = v[e];
          }
        }
      } else {
        const int npool = max(0, nrow) >> 1;  // floor: an odd last row is dropped (max_pool1d)
        const long long obase = tl.row0_out + (tbase >> 1);
#pragma unroll
        for (int p0 = half * 8; p0 < half * 8 + 8; p0 += 4) {
          float4 rv[8], v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            rv[e] = (2 * p0 + e < 2 * npool) ? __ldg(reinterpret_cast<const float4*>(x + (rbase + 2 * p0 + e) * C) + lane)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            v[e] = *reinterpret_cast<const float4*>(et + (2 * p0 + e) * C + ((lane ^ ((2 * p0 + e) & 7)) << 2));
#pragma unroll
          for (int e = 0; e < 8; ++e) { v[e].x += rv[e].x; v[e].y += rv[e].y; v[e].z += rv[e].z; v[e].w += rv[e].w; }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float4 m;
            m.x = fmaxf(v[2 * e].x, v[2 * e + 1].x); m.y = fmaxf(v[2 * e].y, v[2 * e + 1].y);
            m.z = fmaxf(v[2 * e].z, v[2 * e + 1].z); m.w = fmaxf(v[2 * e].w, v[2 * e + 1].w);
            if (relu_final) { m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f); m.z = fmaxf(m.z, 0.f); m.w = fmaxf(m.w, 0.f); }
            if (p0 + e < npool) *(reinterpret_cast<float4*>(out + (obase + p0 + e) * C) + lane) = m;
          }
        }
      }
      // every epilogue warp must be done with the staging bytes before anyone writes the next Y
      if (warp == 4 && lane == 0) MUCON_TR(11, it);  // this warp's rows stored
      named_bar_sync(2, 32 * EPI_WARPS);
      if (warp == 4 && lane == 0) MUCON_TR(10, it);  // tile done
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

// Slab variant for small dilations (dil <= kSlabMaxDil): the three taps of the dilated conv read
// overlapping rows, so per k-block ONE activation slab of 128 + 2*dil rows is loaded and the taps are
// row-shifted SWIZZLE_128B views of it: GEMM 1 takes in 272 KB instead of 384 KB per tile, and with
// eight epilogue warps GEMM 1's operand supply is what the tile time hangs on.  Two rings keep the
// pipeline deep: activation slabs (3 x 20 KB) and weight tiles (6 x 16 KB; the 1x1 weights of GEMM 2
// travel through the same ring).  Everything after the MMAs (epilogues, pooling, stores) is the code
// of wavenet_layer_kernel.
constexpr int kSlabMaxDil = 16;
constexpr int SLAB_BYTES = (BM + 2 * kSlabMaxDil) * 128;   // 20 KB, a multiple of the 1024-byte swizzle atom
constexpr int SLAB_STAGES = 3;
constexpr int W_STAGES = 6;
constexpr int SSMEM_BYTES = 1024 + SLAB_STAGES * SLAB_BYTES + W_STAGES * B_BYTES + Y_BYTES + 512;
static_assert(SLAB_BYTES % 1024 == 0 && B_BYTES % 1024 == 0 && SSMEM_BYTES <= 227 * 1024, "slab kernel shared memory");

__global__ void __launch_bounds__(LTHREADS, 1)
wavenet_layer_slab_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWd,
                     const __grid_constant__ CUtensorMap tmW1, const Tile* __restrict__ tiles, int num_tiles, int dil,
                     const float* __restrict__ bd, const float* __restrict__ b1, const float* __restrict__ x,
                     float* __restrict__ out, int pool, int relu_final) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* slab_mem = base;                              // [SLAB_STAGES][SLAB_BYTES] activation slabs
  unsigned char* w_mem = base + SLAB_STAGES * SLAB_BYTES;      // [W_STAGES][B_BYTES] weight tiles
  unsigned char* ybuf = w_mem + W_STAGES * B_BYTES;            // all three regions are multiples of 1024 bytes
  uint64_t* fullS = reinterpret_cast<uint64_t*>(ybuf + Y_BYTES);
  uint64_t* readyS = fullS + SLAB_STAGES;
  uint64_t* emptyS = readyS + SLAB_STAGES;
  uint64_t* fullW = emptyS + SLAB_STAGES;
  uint64_t* emptyW = fullW + W_STAGES;
  uint64_t* a1full = emptyW + W_STAGES;  // accumulator 1 complete
  uint64_t* a1empty = a1full + 1;       // accumulator 1 drained by the epilogue
  uint64_t* yready = a1empty + 1;       // Y written and published to the async proxy
  uint64_t* a2full = yready + 1;        // accumulator 2 complete (Y no longer read)
  uint64_t* a2empty = a2full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLAB_STAGES; ++s) { mbar_init(&fullS[s], 1); mbar_init(&readyS[s], 1); mbar_init(&emptyS[s], 1); }
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(&fullW[s], 1); mbar_init(&emptyW[s], 1); }
    mbar_init(a1full, 1); mbar_init(a1empty, EPI_WARPS); mbar_init(yready, EPI_WARPS); mbar_init(a2full, 1); mbar_init(a2empty, EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tile walk: alone, CTA b takes tiles b, b + grid, ...; paired, cluster c takes tile pairs c, c + clusters, ...
  const int tile_first = static_cast<int>(blockIdx.x);
  const int tile_step = static_cast<int>(gridDim.x);
  const uint32_t slab_bytes = static_cast<uint32_t>(BM + 2 * dil) * 128u;  // rows of 32 floats

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWd) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
      int ss = 0, ws = 0;
      uint32_t phs = 0, phw = 0;
      for (int ti = tile_first; ti < num_tiles; ti += tile_step) {
        const Tile tl = tiles[ti];
        // GEMM 1: per k-block ONE activation slab of 128 + 2*dil rows (the three taps are row-shifted
        // views of it) into the slab ring, and the three taps' weight tiles into the weight ring
        const int row = static_cast<int>(tl.row0) + tl.t0 - dil;  // may be negative: TMA zero-fills
        for (int kc = 0; kc < KB_PER_TAP; ++kc) {
          mbar_wait(&emptyS[ss], phs ^ 1);
          mbar_arrive_expect_tx(&fullS[ss], slab_bytes);
          tma_load_2d(slab_mem + ss * SLAB_BYTES, &tmX, kc * BK, row, &fullS[ss]);
          if (++ss == SLAB_STAGES) { ss = 0; phs ^= 1; }
          for (int tap = 0; tap < 3; ++tap) {
            mbar_wait(&emptyW[ws], phw ^ 1);
            mbar_arrive_expect_tx(&fullW[ws], B_BYTES);
            tma_load_2d(w_mem + ws * B_BYTES, &tmWd, kc * BK, tap * C, &fullW[ws]);
            if (++ws == W_STAGES) { ws = 0; phw ^= 1; }
          }
        }
        // GEMM 2: the 1x1 weights, through the weight ring
        for (int kc = 0; kc < KB_PER_TAP; ++kc) {
          mbar_wait(&emptyW[ws], phw ^ 1);
          mbar_arrive_expect_tx(&fullW[ws], B_BYTES);
          tma_load_2d(w_mem + ws * B_BYTES, &tmW1, kc * BK, 0, &fullW[ws]);
          if (++ws == W_STAGES) { ws = 0; phw ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ fix-up warp =================================
    int ss = 0;
    uint32_t phs = 0;
    for (int ti = tile_first; ti < num_tiles; ti += tile_step) {
      const Tile tl = tiles[ti];
      // slab row r holds time step t0 - dil + r: rows outside [0, T) are Conv1d's zero padding
      const int rows = BM + 2 * dil;
      const int lo = dil - tl.t0;            // rows below lo are before the video
      const int hi = tl.T - tl.t0 + dil;     // rows from hi on are after it
      const bool fix = lo > 0 || hi < rows;
      for (int kc = 0; kc < KB_PER_TAP; ++kc) {
        mbar_wait(&fullS[ss], phs);
        if (fix) {
          zero_pad_rows(slab_mem + ss * SLAB_BYTES, lo, hi, rows, lane);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&readyS[ss]);
        if (++ss == SLAB_STAGES) { ss = 0; phs ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = instr_desc_tf32(BM, BN);
    int ss = 0, ws = 0, it = 0;
    uint32_t phs = 0, phw = 0;
    const uint32_t d1 = tmem_base, d2 = tmem_base + BN;
    for (int ti = tile_first; ti < num_tiles; ti += tile_step, ++it) {
      const uint32_t tph = it & 1;
      // ---- GEMM 1: dilated conv
      mbar_wait(a1empty, tph ^ 1);
      tc_fence_after();
      for (int kc = 0; kc < KB_PER_TAP; ++kc) {
        mbar_wait(&readyS[ss], phs);
        const uint32_t a_addr = smem_u32(slab_mem + ss * SLAB_BYTES);
        for (int tap = 0; tap < 3; ++tap) {
          mbar_wait(&fullW[ws], phw);
          tc_fence_after();
          if (lane == 0) {
            // tap `tap` reads slab rows tap*dil .. tap*dil + 127: the same SWIZZLE_128B descriptor with its
            // start address moved by whole 128-byte rows (the swizzle is a function of the absolute
            // shared-memory address, so a row shift needs no base-offset field; checked on the device)
            const uint64_t adesc = smem_desc(a_addr + static_cast<uint32_t>(tap * dil) * 128u);
            const uint64_t bdesc = smem_desc(smem_u32(w_mem + ws * B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) mma_tf32(d1, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0);
            mma_commit(&emptyW[ws]);
            if (tap == 2) mma_commit(&emptyS[ss]);
          }
          __syncwarp();
          if (++ws == W_STAGES) { ws = 0; phw ^= 1; }
        }
        if (++ss == SLAB_STAGES) { ss = 0; phs ^= 1; }
      }
      if (lane == 0) mma_commit(a1full);
      __syncwarp();
      // ---- GEMM 2: 1x1 conv on relu(acc1 + bd), which the epilogue warps wrote to ybuf
      mbar_wait(yready, tph);
      mbar_wait(a2empty, tph ^ 1);
      tc_fence_after();
      for (int kc = 0; kc < KB_PER_TAP; ++kc) {
        mbar_wait(&fullW[ws], phw);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = smem_desc(smem_u32(ybuf + kc * A_BYTES));
          const uint64_t bdesc = smem_desc(smem_u32(w_mem + ws * B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) mma_tf32(d2, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
          mma_commit(&emptyW[ws]);
        }
        __syncwarp();
        if (++ws == W_STAGES) { ws = 0; phw ^= 1; }
      }
      if (lane == 0) mma_commit(a2full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    // Eight epilogue warps: the per-tile timeline (clock stamps, scratch build) showed the four epilogue
    // warps busy ~12.5 k of the 14.4 k cycles of a tile (8.5 k of it in the residual-load / store phase)
    // while the MMA and TMA warps waited for them.  Two warps now share a TMEM lane quarter: each takes
    // 64 of the 128 accumulator columns in the TMEM phases and 16 of the quarter's 32 rows in the
    // coalesced phase.
    const int q = warp & 3;            // TMEM lane quarter (a warp may only touch lanes 32*(warp%4) ..)
    const int half = (warp - 4) >> 2;  // which 64 accumulator columns / which 16 rows of the quarter
    const int r = q * 32 + lane;       // tile row owned in the TMEM-load phases
    int it = 0;
    for (int ti = tile_first; ti < num_tiles; ti += tile_step, ++it) {
      const Tile tl = tiles[ti];
      const uint32_t tph = it & 1;
      // ---- epilogue 1: acc1 -> relu(. + bd) -> ybuf in the K-major SWIZZLE_128B operand layout
      mbar_wait(a1full, tph);
      tc_fence_after();
#pragma unroll
      for (int kh = 0; kh < BN / 64; ++kh) {
        const int kc = half * (BN / 64) + kh;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kc * 32, v);
        unsigned char* rowp = ybuf + kc * A_BYTES + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = fmaxf(__uint_as_float(v[4 * j + 0]) + __ldg(bd + kc * 32 + 4 * j + 0), 0.f);
          o.y = fmaxf(__uint_as_float(v[4 * j + 1]) + __ldg(bd + kc * 32 + 4 * j + 1), 0.f);
          o.z = fmaxf(__uint_as_float(v[4 * j + 2]) + __ldg(bd + kc * 32 + 4 * j + 2), 0.f);
          o.w = fmaxf(__uint_as_float(v[4 * j + 3]) + __ldg(bd + kc * 32 + 4 * j + 3), 0.f);
          *reinterpret_cast<float4*>(rowp + ((j ^ (r & 7)) << 4)) = o;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(a1empty); mbar_arrive(yready); }
      // ---- epilogue 2: acc2 + b1 -> staging (the ybuf bytes, free once GEMM 2 has completed)
      mbar_wait(a2full, tph);
      tc_fence_after();
      float* et = reinterpret_cast<float*>(ybuf) + q * (32 * C);  // [32 rows][128], 16-byte chunks XOR-swizzled by row
#pragma unroll
      for (int kh = 0; kh < BN / 64; ++kh) {
        const int kc = half * (BN / 64) + kh;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + BN + kc * 32, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = __uint_as_float(v[4 * j + 0]) + __ldg(b1 + kc * 32 + 4 * j + 0);
          o.y = __uint_as_float(v[4 * j + 1]) + __ldg(b1 + kc * 32 + 4 * j + 1);
          o.z = __uint_as_float(v[4 * j + 2]) + __ldg(b1 + kc * 32 + 4 * j + 2);
          o.w = __uint_as_float(v[4 * j + 3]) + __ldg(b1 + kc * 32 + 4 * j + 3);
          const int chunk = kc * 8 + j;
          *reinterpret_cast<float4*>(et + lane * C + ((chunk ^ (lane & 7)) << 2)) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2empty);
      // both warps of a quarter have staged their 64 columns before either reads whole rows
      named_bar_sync(3 + q, 64);
      // ---- coalesced output: + residual, optional ReLU, optional max-pool of adjacent rows;
      // the quarter's 32 rows are split between its two warps
      const int tbase = tl.t0 + q * 32;
      const int nrow = min(32, tl.T - tbase);
      const long long rbase = tl.row0 + tbase;
      if (!pool) {
#pragma unroll
        for (int r0 = half * 16; r0 < half * 16 + 16; r0 += 8) {
          float4 rv[8], v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            rv[e] = (r0 + e < nrow) ? __ldg(reinterpret_cast<const float4*>(x + (rbase + r0 + e) * C) + lane)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            v[e] = *reinterpret_cast<const float4*>(et + (r0 + e) * C + ((lane ^ ((r0 + e) & 7)) << 2));
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v[e].x += rv[e].x; v[e].y += rv[e].y; v[e].z += rv[e].z; v[e].w += rv[e].w;
            if (relu_final) {
              v[e].x = fmaxf(v[e].x, 0.f); v[e].y = fmaxf(v[e].y, 0.f); v[e].z = fmaxf(v[e].z, 0.f); v[e].w = fmaxf(v[e].w, 0.f);
            }
            if (r0 + e < nrow) *(reinterpret_cast<float4*>(out + (tl.row0_out + tbase + r0 + e) * C) + lane) = v[e];
          }
        }
      } else {
        const int npool = max(0, nrow) >> 1;  // floor: an odd last row is dropped (max_pool1d)
        const long long obase = tl.row0_out + (tbase >> 1);
#pragma unroll
        for (int p0 = half * 8; p0 < half * 8 + 8; p0 += 4) {
          float4 rv[8], v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            rv[e] = (2 * p0 + e < 2 * npool) ? __ldg(reinterpret_cast<const float4*>(x + (rbase + 2 * p0 + e) * C) + lane)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            v[e] = *reinterpret_cast<const float4*>(et + (2 * p0 + e) * C + ((lane ^ ((2 * p0 + e) & 7)) << 2));
#pragma unroll
          for (int e = 0; e < 8; ++e) { v[e].x += rv[e].x; v[e].y += rv[e].y; v[e].z += rv[e].z; v[e].w += rv[e].w; }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float4 m;
            m.x = fmaxf(v[2 * e].x, v[2 * e + 1].x); m.y = fmaxf(v[2 * e].y, v[2 * e + 1].y);
            m.z = fmaxf(v[2 * e].z, v[2 * e + 1].z); m.w = fmaxf(v[2 * e].w, v[2 * e + 1].w);
            if (relu_final) { m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f); m.z = fmaxf(m.z, 0.f); m.w = fmaxf(m.w, 0.f); }
            if (p0 + e < npool) *(reinterpret_cast<float4*>(out + (obase + p0 + e) * C) + lane) = m;
          }
        }
      }
      // every epilogue warp must be done with the staging bytes before anyone writes the next Y
      named_bar_sync(2, 32 * EPI_WARPS);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

}  // namespace layer
}  // namespace mucon
