// backbone_train.cuh -- weight gradients of the backbone's convolutions on tcgen05 (training step, config c5).
//
// Reference: the backward pass autograd derives for src/core/modules/temporal.py:43-53,128-147 when
// src/mucon/trainers.py:125-131 calls loss.backward().  For a convolution  y[t] = sum_tap W[tap] x[t + s_tap] + b
// over time-major rows (channels contiguous, videos concatenated, zero padding at a video's ends):
//
//   dW[tap][co][ci] = sum_t dY[t, co] * X[t + s_tap, ci]          (t and t + s_tap inside the video)
//   db[co]          = sum_t dY[t, co]
//
// i.e. a GEMM whose K dimension is TIME.  Both operands are stored with the M/N dimension (channels) contiguous,
// so they are fed to `tcgen05.mma.kind::tf32` as MN-major SWIZZLE_128B operands: a TMA box of [32 channels x KT
// rows] lands as KT 128-byte rows (32-byte-atom swizzle, 4-row atoms; 8 rows = 8 values of K = one MMA), four such
// boxes side by side (leading byte offset = box size) give the 128 channels of M (or N).  No transposed copy of the activations.
//
// One launch handles one (dY, X) pair and up to kMaxJobs "jobs" -- a job is (row shift, column block of X,
// output offset): the three taps of a dilated convolution, the single tap of a 1x1, or the sixteen 128-column
// blocks of the 2048-wide features for the input projection.  A CTA owns a contiguous range of 128-row time tiles
// and up to four jobs (blockIdx.y picks the group of four): the dY tile of a stage is loaded once and multiplied
// against the X tile of every live job, each job accumulating in its own 128 TMEM columns over the CTA's whole
// range; at the end the accumulators are added to the gradient buffer with vector reductions (red.global.add.v4.f32).
// The bias gradient is the column sum of the dY tiles, taken by the fix-up warp from the tiles already in shared
// memory (no extra pass over dY).
// 224 threads: warp 0 TMA producer, warp 1 TMEM + MMA issuer, warp 2 fix-up (zero padding, bias gradient),
// warps 3-6 epilogue.
#pragma once
#include <cuda.h>

#include "backbone_gemm.cuh"

namespace mucon {
namespace wgrad {

using namespace gemm;
constexpr int WTHREADS = 224;
constexpr int C = 128;
constexpr int KT = 32;                      // rows (K) per stage
constexpr int OP_BYTES = KT * C * 4;        // one operand tile: four [KT x 32] boxes = 16 KB
constexpr int BOX_BYTES = KT * 32 * 4;      // 4 KB
constexpr int kJobsPerCta = 4;              // 4 x 128 TMEM columns
constexpr int kMaxJobs = 16;
constexpr int kMaxStages = 6;
constexpr int WSMEM_MAX = 200 * 1024;

struct Job {
  int shift;     // X row = dY row + shift
  int xblk;      // first 32-column block of X (column offset / 32)
  long long out_off;  // element offset of this job's [128 x 128] block in the gradient buffer
};
struct Jobs {
  int n;
  int ldo;       // row pitch (elements) of the output blocks
  Job j[kMaxJobs];
};

// MN-major shared-memory matrix descriptor for 32-bit operands.  TF32 operands can only be read MN-major from the
// "128-byte swizzle with 32-byte atoms" layout (cute: Layout_MN_SW128_32B_Atom = Swizzle<2,5,2>, descriptor layout
// type 1 = SWIZZLE_128B_BASE32B; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 128-byte rows (32 channels) at a 128-byte
// pitch along K, the 32-byte chunks of a row XOR-permuted by (row & 3), atoms of 4 rows.  Canonical layout in 16-byte
// units ((8,n),(4,k)):((1,LBO),(8,SBO)): LBO = byte distance between 32-channel boxes, SBO = byte distance between
// groups of 4 K-rows (512).
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;
  return d;
}
// instruction descriptor as gemm::instr_desc_tf32 with A and B MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t instr_desc_tf32_mn(int M, int N) {
  return instr_desc_tf32(M, N) | (1u << 15) | (1u << 16);
}

// one operand tile = four [KT rows x 32 channels] boxes side by side (2-D tiled loads, 128-byte swizzle)
__device__ __forceinline__ void load_operand(unsigned char* dst, const CUtensorMap* map, int col0, int row, uint64_t* bar) {
#pragma unroll
  for (int b = 0; b < 4; ++b) tma_load_2d(dst + b * BOX_BYTES, map, col0 + 32 * b, row, bar);
}

__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// rows [t_lo, t_hi) of the video are in this stage; is any of them valid for a shift?
__device__ __forceinline__ bool job_live(int shift, int t_lo, int t_hi, int T) {
  return t_lo + shift < T && t_hi - 1 + shift >= 0;
}

__global__ void __launch_bounds__(WTHREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
             const convgemm::Tile* __restrict__ tiles, int num_tiles, const Jobs jobs, int stages,
             float* __restrict__ dW, float* __restrict__ dbias) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int j0 = blockIdx.y * kJobsPerCta;
  const int nj = min(kJobsPerCta, jobs.n - j0);
  const int stage_bytes = (1 + nj) * OP_BYTES;
  unsigned char* stage_mem = base;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + stages * stage_bytes);
  uint64_t* ready = full + kMaxStages;
  uint64_t* empty = ready + kMaxStages;
  uint64_t* tfull = empty + kMaxStages;    // accumulators complete
  uint64_t* tzero = tfull + 1;             // accumulators zeroed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tzero + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    mbar_init(tzero, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's contiguous range of time tiles
  const int ti_lo = static_cast<int>(static_cast<long long>(num_tiles) * blockIdx.x / gridDim.x);
  const int ti_hi = static_cast<int>(static_cast<long long>(num_tiles) * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDY) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int ti = ti_lo; ti < ti_hi; ++ti) {
        const convgemm::Tile tl = tiles[ti];
        for (int kb = 0; kb < BM / KT; ++kb) {
          const int t_lo = tl.t0 + kb * KT;
          if (t_lo >= tl.T) break;
          const int t_hi = min(t_lo + KT, tl.T);
          int live = 0;
          for (int j = 0; j < nj; ++j) live += job_live(jobs.j[j0 + j].shift, t_lo, t_hi, tl.T);
          if (!live) continue;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], (1 + live) * OP_BYTES);
          unsigned char* st = stage_mem + s * stage_bytes;
          const int row = static_cast<int>(tl.row0) + t_lo;
          load_operand(st, &tmDY, 0, row, &full[s]);
          for (int j = 0; j < nj; ++j) {
            const Job jb = jobs.j[j0 + j];
            if (!job_live(jb.shift, t_lo, t_hi, tl.T)) continue;
            load_operand(st + (1 + j) * OP_BYTES, &tmX, jb.xblk * 32, row + jb.shift, &full[s]);  // rows < 0: zero fill
          }
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ fix-up warp =================================
    // zero the X rows whose frame (or whose dY frame) lies outside the video; column sums of dY for the bias
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool want_bias = dbias != nullptr && blockIdx.y == 0;
    int s = 0;
    uint32_t ph = 0;
    for (int ti = ti_lo; ti < ti_hi; ++ti) {
      const convgemm::Tile tl = tiles[ti];
      for (int kb = 0; kb < BM / KT; ++kb) {
        const int t_lo = tl.t0 + kb * KT;
        if (t_lo >= tl.T) break;
        const int t_hi = min(t_lo + KT, tl.T);
        int live = 0;
        for (int j = 0; j < nj; ++j) live += job_live(jobs.j[j0 + j].shift, t_lo, t_hi, tl.T);
        if (!live) continue;  // (the host requires a shift-0 job in group 0 when dbias is requested: always live)
        mbar_wait(&full[s], ph);
        unsigned char* st = stage_mem + s * stage_bytes;
        const int nvalid = t_hi - t_lo;  // rows kk >= nvalid of the dY tile belong to the next video
        bool wrote = false;
        for (int j = 0; j < nj; ++j) {
          const int sh = jobs.j[j0 + j].shift;
          if (!job_live(sh, t_lo, t_hi, tl.T)) continue;
          // valid kk: kk < nvalid, t_lo + kk + sh >= 0, t_lo + kk + sh < T
          const int lo = max(0, -(t_lo + sh));
          const int hi = min(nvalid, tl.T - (t_lo + sh));
          if (lo > 0 || hi < KT) {
            float4* b4 = reinterpret_cast<float4*>(st + (1 + j) * OP_BYTES);
            for (int kk = 0; kk < KT; ++kk) {
              if (kk >= lo && kk < hi) continue;
              // row kk of box (lane >> 3), 16-byte chunk (lane & 7): the whole 128-byte row of each box
              b4[(lane >> 3) * (BOX_BYTES / 16) + kk * 8 + (lane & 7)] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            wrote = true;
          }
        }
        if (want_bias) {
          // dY tile: box b holds channels 32b .. 32b+31; row kk at kk * 128 B, 32-byte chunks XOR-swizzled by (kk & 3)
          const float* a = reinterpret_cast<const float*>(st);
          for (int kk = 0; kk < nvalid; ++kk) {
            const int off = kk * 32 + ((((lane >> 3) ^ (kk & 3)) << 3) | (lane & 7));
#pragma unroll
            for (int b = 0; b < 4; ++b) bsum[b] += a[b * (BOX_BYTES / 4) + off];
          }
        }
        if (wrote) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    if (want_bias) {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (bsum[b] != 0.f) atomicAdd(dbias + b * 32 + lane, bsum[b]);
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = instr_desc_tf32_mn(BM, BN);
    mbar_wait(tzero, 0);   // the epilogue warps have zeroed the accumulators
    tc_fence_after();
    int s = 0;
    uint32_t ph = 0;
    for (int ti = ti_lo; ti < ti_hi; ++ti) {
      const convgemm::Tile tl = tiles[ti];
      for (int kb = 0; kb < BM / KT; ++kb) {
        const int t_lo = tl.t0 + kb * KT;
        if (t_lo >= tl.T) break;
        const int t_hi = min(t_lo + KT, tl.T);
        int live = 0;
        for (int j = 0; j < nj; ++j) live += job_live(jobs.j[j0 + j].shift, t_lo, t_hi, tl.T);
        if (!live) continue;
        mbar_wait(&ready[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(stage_mem + s * stage_bytes);
          for (int j = 0; j < nj; ++j) {
            if (!job_live(jobs.j[j0 + j].shift, t_lo, t_hi, tl.T)) continue;
            const uint32_t b_addr = a_addr + (1 + j) * OP_BYTES;
#pragma unroll
            for (int k = 0; k < KT / UMMA_K; ++k) {
              // 8 rows of K = two 512-byte swizzle atoms
              const uint64_t adesc = smem_desc_mn(a_addr + k * 1024, BOX_BYTES, 512);
              const uint64_t bdesc = smem_desc_mn(b_addr + k * 1024, BOX_BYTES, 512);
              mma_tf32(tmem_base + j * BN, adesc, bdesc, idesc, 1u);
            }
          }
          mma_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    if (lane == 0) mma_commit(tfull);
    __syncwarp();
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c = 0; c < nj * (BN / 32); ++c) tmem_st32_zero(lane_base + c * 32);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tzero);
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int co = q * 32 + lane;
    for (int j = 0; j < nj; ++j) {
      float* orow = dW + jobs.j[j0 + j].out_off + static_cast<long long>(co) * jobs.ldo;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_base + j * BN + c * 32, r);
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float a = __uint_as_float(r[e]), b = __uint_as_float(r[e + 1]), cc = __uint_as_float(r[e + 2]),
                      d = __uint_as_float(r[e + 3]);
          if (a != 0.f || b != 0.f || cc != 0.f || d != 0.f) red_add_v4(orow + c * 32 + e, a, b, cc, d);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// max_pool1d(kernel_size = 2) backward on time-major rows: the gradient of pooled row t goes to input row 2t if
// x[2t] >= x[2t+1] (torch's first-maximum rule) else to row 2t+1; a trailing odd row gets zero.
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const int64_t* __restrict__ off_in,
                                                           const int64_t* __restrict__ off_out, int Cc,
                                                           float* __restrict__ dx) {
  const int v = blockIdx.y;
  const int64_t i0 = off_in[v], o0 = off_out[v];
  const int Tin = static_cast<int>(off_in[v + 1] - i0);
  const int To = static_cast<int>(off_out[v + 1] - o0);
  const int c4n = Cc / 4;
  const int64_t n = static_cast<int64_t>(To) * c4n;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / c4n), c = static_cast<int>(i - static_cast<int64_t>(t) * c4n) * 4;
    const float4 a = *reinterpret_cast<const float4*>(x + (i0 + 2 * t) * Cc + c);
    const float4 b = *reinterpret_cast<const float4*>(x + (i0 + 2 * t + 1) * Cc + c);
    const float4 g = *reinterpret_cast<const float4*>(dy + (o0 + t) * Cc + c);
    float4 ga, gb;
    ga.x = a.x >= b.x ? g.x : 0.f; gb.x = a.x >= b.x ? 0.f : g.x;
    ga.y = a.y >= b.y ? g.y : 0.f; gb.y = a.y >= b.y ? 0.f : g.y;
    ga.z = a.z >= b.z ? g.z : 0.f; gb.z = a.z >= b.z ? 0.f : g.z;
    ga.w = a.w >= b.w ? g.w : 0.f; gb.w = a.w >= b.w ? 0.f : g.w;
    *reinterpret_cast<float4*>(dx + (i0 + 2 * t) * Cc + c) = ga;
    *reinterpret_cast<float4*>(dx + (i0 + 2 * t + 1) * Cc + c) = gb;
  }
  if ((Tin & 1) && blockIdx.x == 0)
    for (int c = threadIdx.x; c < Cc; c += blockDim.x) dx[(i0 + Tin - 1) * Cc + c] = 0.f;
}

}  // namespace wgrad
}  // namespace mucon
