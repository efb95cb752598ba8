// backbone_bf16.cuh -- one whole WaveNet layer per launch on tcgen05 `kind::f16` with bf16 operands.
//
//   y = relu(conv_k3_dil(x) + bd);  out = conv_1x1(y) + b1 + x;  [max_pool1d(2)]  [relu]
//   (reference src/core/modules/temporal.py:43-53, pooling :137-139, the ReLU of :144)
//
// What changed against the TF32 layer kernels of backbone_gemm.cuh (which moved 448 KB of operand tiles,
// 57 % of them re-read weights, through the SM per 128-row tile and were bound by that L2 -> SM stream):
//   * activations travel as bf16 (256 B per time step): half the HBM and L2 bytes of every layer;
//   * the layer's weights (3 taps + the 1x1: 4 x 128 x 128 bf16 = 128 KB) are loaded ONCE per CTA and stay
//     resident in shared memory in the K-major SWIZZLE_128B operand layout;
//   * per tile a single activation slab of 128 + 2*dil rows comes in (two 64-channel k-blocks); the three
//     taps are row-shifted descriptor views of it;
//   * nothing but the slab, the weights and a 2 KB identity lives in shared memory.  The residual is added BY THE
//     TENSOR CORE: after GEMM 1 the MMA warp issues eight M128 N16 K16 instructions acc2[:, 16k:16k+16] =
//     X_centre[:, 16k:16k+16] . I (exact: every product is x * 1 or x * 0, accumulated in fp32) and GEMM 2
//     accumulates on top, so the epilogue warps -- the bound of this kernel -- never touch the residual and the
//     slab is released by the MMA commit; relu(acc1 + bd) is rounded to 16 bits and stored back IN PLACE over
//     the first 64 columns of accumulator 1, from where GEMM 2 takes it as its A operand (tcgen05.mma with A in
//     tensor memory); the second epilogue is TMEM -> + b1 -> (ReLU) (max over adjacent rows) -> 16 bits ->
//     32-byte global stores.  32-48 KB in and 16-32 KB out per tile instead of 576 KB;
//   * accumulators are double-buffered in TMEM (2 x 128 + 2 x 128 columns); the MMA warp issues GEMM 1 of tile
//     i+1 before GEMM 2 of tile i and the epilogue warps run epilogue 1 of tile i+1 before epilogue 2 of tile
//     i, so the tensor pipe and the epilogue warps both stay busy.
// Dilations above kMaxSlabDil use the same program with each live tap as a separate 128-row tile in a ring of
// three 32 KB slots: side-tap slots are released by the MMA warp's commit as soon as that tap's instructions have
// retired, the centre tap (issued last) by the commit that follows the residual MMAs, so the loads of
// the next tile overlap this tile's GEMMs (taps that only see padding -- dilation >= video length -- are never
// loaded: layers 8-10 at Breakfast lengths are three-deep rings of centre tiles).
// 384 threads: warp 0 TMA producer, warp 1 MMA + TMEM, warp 2 padding fix-up, warp 3 idle, warps 4-11 epilogue.
#pragma once
#include <cuda.h>

#include "backbone_gemm.cuh"

namespace mucon {
namespace layer16 {

using namespace gemm;  // TMA / mbarrier / tcgen05 helpers, BM = BN = 128

constexpr int C = 128;
constexpr int NKB = 2;                    // k-blocks of 64 bf16 = one 128-byte swizzle row each
constexpr int WTILE = BN * 128;           // one [128 x 64] bf16 weight tile: 16 KB
constexpr int W_BYTES = 8 * WTILE;        // 3 taps x 2 k-blocks + the 1x1's 2 k-blocks: 128 KB
constexpr int LTHREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr int kMaxSlabDil = 32;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int BAR_BYTES = 256;
constexpr int ID_BYTES = 16 * 128;        // a [16 x 16] 16-bit identity in 128-byte K-major SWIZZLE_128B rows: B operand of the residual MMAs

// shared-memory plan of a launch (host and device agree through these)
__host__ __device__ inline int stage_rows(int dil, int slab) { return slab ? BM + 2 * dil : 3 * BM; }
__host__ __device__ inline int kb_bytes_of(int dil, int slab) { return ((stage_rows(dil, slab) + 7) & ~7) * 128; }
__host__ __device__ inline int num_stages(int slab) { return slab ? 2 : 1; }  // x stage bytes = the ring's bytes
constexpr int kRingSlots = 3;               // ring mode (separate taps): slots of one [128 x 128] tap tile (32 KB) each
constexpr int kTapBytes = NKB * BM * 128;
// rows of the (bf16) output tile that are staged in shared memory and leave through a TMA store; the rest of the
// tile (and every tile that crosses the end of its video, and fp32 output) is stored straight from registers
__host__ __device__ inline int staged_rows(int dil, int slab, int pool, int out_f32) {
  if (out_f32) return 0;
#ifdef MUCON_L16_NOSTAGE
  return 0;   // experiment: every tile stored straight from registers (scripts/ab_layers.py)
#endif
  const int free_bytes = SMEM_LIMIT - W_BYTES - ID_BYTES - num_stages(slab) * NKB * kb_bytes_of(dil, slab) - BAR_BYTES;
  int s = (free_bytes / 256) & ~7;
  const int want = pool ? BM / 2 : BM;
  if (s > want) s = want;
  return s < 8 ? 0 : s;
}
__host__ __device__ inline int smem_bytes_of(int dil, int slab, int pool, int out_f32) {
  return W_BYTES + ID_BYTES + num_stages(slab) * NKB * kb_bytes_of(dil, slab) + staged_rows(dil, slab, pool, out_f32) * 256 +
         BAR_BYTES;
}

// both bias vectors travel as a kernel parameter: the epilogue adds them as constant-bank operands
struct BiasPack {
  float bd[C];
  float b1[C];
};

// developer builds (-DMUCON_LAYER_TRACE, scripts/trace_layer16.py): clock64() stamps per role and tile of CTA 0
#ifdef MUCON_LAYER_TRACE
#define MUCON_TR16(ev, it) do { if (blockIdx.x == 0 && (it) < 128) layer::g_trace[ev][it] = clock64(); } while (0)
#else
#define MUCON_TR16(ev, it) do { } while (0)
#endif

struct Tile {
  long long row0;      // first row of the video at the input resolution
  long long row0_out;  // first row of the video in the output tensor (differs when pooling)
  int t0;              // first time step of the tile within the video
  int T;               // video length at the input resolution
};

__device__ __forceinline__ bool tap_live(int shift, int T) { return shift < T && -shift < T; }

// instruction descriptor: D = F32, A = B = BF16 (format 1), both K-major
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N, bool f16 = false) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: A is a [128 x 16] bf16 block in tensor memory, row m in lane m, elements
// 2c / 2c+1 of a row in the low / high half of 32-bit column c
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// two independent IEEE fp32 additions in one instruction (Blackwell packed fp32)
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\t"
      "mov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// The 16-bit activation / weight type of a launch: bf16 (8-bit mantissa, fp32 range) or fp16 (11-bit mantissa:
// rounding the residual stream costs 8x less error, at the price of saturating at +-65504).
template <bool F16>
struct Half2 {
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r;
    if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
    uint32_t r;
    if (F16) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ float lo(uint32_t p) {
    if (!F16) return bf16_lo(p);
    float f;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}" : "=f"(f) : "r"(p));
    return f;
  }
  static __device__ __forceinline__ float hi(uint32_t p) {
    if (!F16) return bf16_hi(p);
    float f;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, h;\n\t}" : "=f"(f) : "r"(p));
    return f;
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    uint32_t r;
    if (F16) asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
  }
};
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

struct EpiCtx {
  const Tile* tiles;
  unsigned char* stage_mem;
  unsigned char* staging;
  uint64_t *a1full, *yready, *a2full, *a2free;
  const CUtensorMap* tmO;
  void* out;
  uint32_t tmem_base;
  int n_my, nslot, LA, unit_bytes, kb_bytes, crow0, slab, dil, S, pool, relu_final, out_f32;
};

// The eight epilogue warps (H = which 64 accumulator columns / which k-block of the slab this warp owns).
template <int H, bool F16>
__device__ __forceinline__ void epilogue_warps(const EpiCtx& c, const BiasPack& bias) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;            // TMEM lane quarter (a warp may only touch lanes 32*(warp%4) ..)
  const int r = q * 32 + lane;       // tile row owned by this thread
  const uint32_t lane_base = c.tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const bool leader = warp == 4 && lane == 0;
  bool store_pending = false;
  for (int i = 0; i < c.n_my + c.LA; ++i) {
    if (i < c.n_my) {
      // ---- epilogue 1 of tile i
      const int acc = i & 1;
      mbar_wait(&c.a1full[acc], (i >> 1) & 1);
      tc_fence_after();
      if (leader) MUCON_TR16(6, i);
      // (the residual x is added by the tensor core: the MMA warp multiplies the tile's centre rows with an identity
      // into accumulator 2 -- see the MMA issuer; b1 is added in epilogue 2)
      // (b) acc1 -> relu(. + bd) -> bf16 -> back over accumulator 1 (columns 32H .. 32H+31): A operand of GEMM 2
      {
        uint32_t v0[32], v1[32];
        tmem_ld32(lane_base + acc * BN + H * 64, v0);
        tmem_ld32(lane_base + acc * BN + H * 64 + 32, v1);
        if (leader) MUCON_TR16(11, i);
        // the other warp of this quarter reads columns that this one overwrites (and vice versa)
        named_bar_sync(3 + q, 64);
        if (leader) MUCON_TR16(14, i);
        uint32_t y[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          y[e] = Half2<F16>::pack_relu(__uint_as_float(v0[2 * e]) + bias.bd[H * 64 + 2 * e],
                                __uint_as_float(v0[2 * e + 1]) + bias.bd[H * 64 + 2 * e + 1]);
          y[16 + e] = Half2<F16>::pack_relu(__uint_as_float(v1[2 * e]) + bias.bd[H * 64 + 32 + 2 * e],
                                     __uint_as_float(v1[2 * e + 1]) + bias.bd[H * 64 + 32 + 2 * e + 1]);
        }
        tmem_st32(lane_base + acc * BN + H * 32, y);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&c.yready[acc]);
      if (leader) MUCON_TR16(7, i);
    }
    if (i >= c.LA) {
      // ---- epilogue 2 of tile j: acc2 (-> ReLU) (-> max over adjacent rows) -> bf16 -> shared-memory staging ->
      // TMA store (rows beyond the staging area, tiles that cross the video's end and fp32 output: 32-byte stores
      // straight from registers)
      const int j = i - c.LA;
      const Tile tl = c.tiles[blockIdx.x + j * gridDim.x];
      const int acc = j & 1;
      mbar_wait(&c.a2full[acc], (j >> 1) & 1);
      tc_fence_after();
      if (leader) MUCON_TR16(8, j);
      const int t = tl.t0 + r;
      const bool row_ok = t < tl.T;
      const bool pair_ok = (t | 1) < tl.T;  // both rows of the pooling pair inside the video (floor)
      const bool staged = c.S > 0 && tl.t0 + BM <= tl.T;  // CTA-uniform
      if (staged) {
        // the previous TMA store must have read the staging area before it is overwritten
        if (leader && store_pending) tma_store_wait_read();
        named_bar_sync(2, 32 * EPI_WARPS);
      }
      const int orow = c.pool ? (r >> 1) : r;  // row of the output tile this thread (pair) produces
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t v[32];
        tmem_ld32(lane_base + 2 * BN + acc * BN + H * 64 + c2 * 32, v);
        const int col = H * 64 + c2 * 32;
        if (c2 == 1) {   // both halves of this warp's accumulator-2 columns are in registers: the MMA warp may reuse it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&c.a2free[acc]);
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + bias.b1[col + e]);
        if (c2 == 0 && leader) MUCON_TR16(15, j);
        if (c2 == 1 && leader) MUCON_TR16(16, j);
        if (c.out_f32) {
          // (the host never asks for pooling together with fp32 output)
          if (c.relu_final) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(fmaxf(__uint_as_float(v[e]), 0.f));
          }
          if (row_ok) {
            float* op = reinterpret_cast<float*>(c.out) + (tl.row0_out + t) * C + col;
#pragma unroll
            for (int g = 0; g < 4; ++g) st_global_v8(op + 8 * g, v + 8 * g);
          }
        } else {
          uint32_t p[16];
          if (c.relu_final) {
#pragma unroll
            for (int e = 0; e < 16; ++e) p[e] = Half2<F16>::pack_relu(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) p[e] = Half2<F16>::pack(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          }
          unsigned char* srow = c.staging + H * (c.S * 128) + orow * 128;  // k-block half H, [S rows x 128 B], SWIZZLE_128B
          if (!c.pool) {
            if (staged && orow < c.S) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                *reinterpret_cast<uint4*>(srow + (((c2 * 4 + g) ^ (orow & 7)) << 4)) =
                    make_uint4(p[4 * g], p[4 * g + 1], p[4 * g + 2], p[4 * g + 3]);
            } else if (row_ok) {
              unsigned short* op = reinterpret_cast<unsigned short*>(c.out) + (tl.row0_out + t) * C + col;
              st_global_v8(op, p);
              st_global_v8(op + 16, p + 8);
            }
          } else {
            // rounding to bf16 is monotonic: max of the rounded values = rounded max.  Adjacent rows are adjacent lanes.
            uint32_t m[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) m[e] = Half2<F16>::max2(p[e], __shfl_xor_sync(0xffffffffu, p[e], 1));
            uint32_t w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) w[e] = (lane & 1) ? m[8 + e] : m[e];  // even lane: columns 0-15, odd lane: 16-31
            if (staged && orow < c.S) {
              const int ch = c2 * 4 + ((lane & 1) << 1);
              *reinterpret_cast<uint4*>(srow + ((ch ^ (orow & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(srow + (((ch + 1) ^ (orow & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
            } else if (pair_ok) {
              unsigned short* op = reinterpret_cast<unsigned short*>(c.out) + (tl.row0_out + (t >> 1)) * C + col + ((lane & 1) << 4);
              st_global_v8(op, w);
            }
          }
        }
      }
      if (leader) MUCON_TR16(17, j);
      if (staged) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> the TMA's (async) proxy
        if (leader) MUCON_TR16(18, j);
        named_bar_sync(2, 32 * EPI_WARPS);
        if (leader) MUCON_TR16(19, j);
        if (leader) {
          const int orow0 = static_cast<int>(tl.row0_out) + (c.pool ? (tl.t0 >> 1) : tl.t0);
          tma_store_2d(c.tmO, c.staging, 0, orow0);
          tma_store_2d(c.tmO, c.staging + c.S * 128, 64, orow0);
          tma_store_commit();
        }
        store_pending = true;
      }
      if (leader) MUCON_TR16(9, j);
    }
  }
  if (leader && store_pending) tma_store_wait_all();  // the stores are complete before the CTA exits
}

template <bool F16>
__global__ void __launch_bounds__(LTHREADS, 1)
wavenet_layer_bf16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWd,
                          const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmO,
                          const __grid_constant__ BiasPack bias, const Tile* __restrict__ tiles, int num_tiles,
                          int dil, int slab, void* __restrict__ out, int pool, int relu_final, int out_f32) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw;  // the launch asks for exactly what it uses: no slack for re-alignment
  if ((smem_u32(base) & 1023u) != 0) __trap();
  // stage geometry: a stage holds NKB k-blocks of R rows x 128 bytes.  Slab mode: R = 128 + 2*dil rows starting at
  // time step t0 - dil, tap `tap` is the 128 rows from row tap*dil.  Otherwise three separate 128-row tap tiles.
  // Unit geometry.  Slab mode: a unit is a slab of NKB k-blocks of R = 128 + 2*dil rows x 128 bytes starting at time
  // step t0 - dil (tap `tap` = the 128 rows from row tap*dil), one unit per tile, two slots.  Ring mode: a unit is
  // one live tap's [128 rows] x NKB k-blocks tile, up to three units per tile (side taps first, the centre last),
  // three slots.
  const int R = slab ? BM + 2 * dil : BM;
  const int kb_bytes = slab ? kb_bytes_of(dil, 1) : BM * 128;
  const int unit_bytes = NKB * kb_bytes;
  const int nslot = slab ? 2 : kRingSlots;
  const int LA = 1;                        // GEMM 1 / epilogue 1 run one tile ahead of GEMM 2 / epilogue 2
  // out_f32 carries two flags: bit 0 = fp32 output, bit 1 = NO residual (out = conv_1x1(relu(conv(x) + bd)) + b1: with
  // an identity centre tap and dead side taps this is a plain 1x1 conv of non-negative 16-bit rows -- last_conv)
  const bool no_res = (out_f32 & 2) != 0;
  out_f32 &= 1;
  const int S = staged_rows(dil, slab, pool, out_f32);
  unsigned char* w_mem = base;                       // [8][WTILE]: (tap, kb) tiles of the dilated conv, then the 1x1's
  unsigned char* id_mem = base + W_BYTES;            // [16 x 16] identity (16-bit, K-major SWIZZLE_128B rows)
  unsigned char* stage_mem = base + W_BYTES + ID_BYTES;  // the unit ring
  unsigned char* staging = stage_mem + num_stages(slab) * NKB * kb_bytes_of(dil, slab);  // [2][S rows x 128 B], SWIZZLE_128B
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + S * 256);
  uint64_t* wfull = bars;          // weights resident
  uint64_t* fullS = bars + 1;      // [3] unit landed (TMA bytes)
  uint64_t* readyS = bars + 4;     // [3] unit padded (fix-up warp)
  uint64_t* emptyM = bars + 10;    // [3] unit released by the MMA warp's commit (its last reader is an MMA)
  uint64_t* a1full = bars + 13;    // [2] accumulator 1 complete
  uint64_t* a1free = bars + 15;    // [2] GEMM 2 retired: accumulator 1 / Y may be overwritten
  uint64_t* yready = bars + 17;    // [2] Y (16-bit, over accumulator 1) stored
  uint64_t* a2full = bars + 19;    // [2] accumulator 2 complete
  uint64_t* a2free = bars + 21;    // [2] accumulator 2 drained by epilogue 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int s = 0; s < 3; ++s) {
      mbar_init(&fullS[s], 1); mbar_init(&readyS[s], 1); mbar_init(&emptyM[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a1full[s], 1); mbar_init(&a1free[s], 1); mbar_init(&yready[s], EPI_WARPS); mbar_init(&a2full[s], 1);
      mbar_init(&a2free[s], EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // all of TMEM: accumulators 1 at columns 0 / 128, accumulators 2 at 256 / 384
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the identity the residual is multiplied with: element (n, k = n) = 1 in the K-major SWIZZLE_128B operand layout
  // (row n = 128 bytes of 64 k values, 16-byte chunks XOR-ed with n & 7)
  for (int i = threadIdx.x; i < ID_BYTES / 16; i += LTHREADS) reinterpret_cast<uint4*>(id_mem)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (threadIdx.x < 16) {
    const int n = threadIdx.x;
    *reinterpret_cast<unsigned short*>(id_mem + n * 128 + (((n >> 3) ^ (n & 7)) << 4) + (n & 7) * 2) =
        F16 ? static_cast<unsigned short>(0x3C00) : static_cast<unsigned short>(0x3F80);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_my = (static_cast<int>(blockIdx.x) < num_tiles)
                       ? (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;
  // The units of a tile, in the order they are loaded and consumed, packed two bits per unit (no local arrays in
  // the single-thread roles): tap of unit k = (code >> 2k) & 3 (slab mode: one unit, tap code 3 = "the slab").
  auto tile_units = [&](const Tile& tl, uint32_t& code) {
    if (slab) { code = 3u; return 1; }
    int n = 0;
    code = 0u;
    if (tap_live(-dil, tl.T)) { code |= 0u << (2 * n); ++n; }
    if (tap_live(dil, tl.T)) { code |= 2u << (2 * n); ++n; }
    code |= 1u << (2 * n);  // the centre tap is always live and comes last: its slot is the one the epilogue releases
    return n + 1;
  };
  // does the unit need rows zeroed (Conv1d padding at the video's ends)?  Only then does the MMA warp wait for the
  // fix-up warp; otherwise it goes straight from the TMA's barrier
  auto needs_fix = [&](const Tile& tl, int tap) {
    if (tap == 3) return dil - tl.t0 > 0 || tl.T - tl.t0 + dil < R;
    const int start = tl.t0 + (tap - 1) * dil;
    return start < 0 || start + BM > tl.T;
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWd) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
      if (n_my > 0) {
        mbar_arrive_expect_tx(wfull, W_BYTES);
        for (int tap = 0; tap < 3; ++tap)
          for (int kb = 0; kb < NKB; ++kb) tma_load_2d(w_mem + (tap * NKB + kb) * WTILE, &tmWd, kb * 64, tap * C, wfull);
        for (int kb = 0; kb < NKB; ++kb) tma_load_2d(w_mem + (3 * NKB + kb) * WTILE, &tmW1, kb * 64, 0, wfull);
      }
      int s = 0;
      uint32_t used = 0, phM = 0;  // per slot: used before / phase of its release barrier
      for (int i = 0; i < n_my; ++i) {
        const Tile tl = tiles[blockIdx.x + i * gridDim.x];
        uint32_t code;
        const int nu = tile_units(tl, code);
        for (int k = 0; k < nu; ++k) {
          const int tap = static_cast<int>((code >> (2 * k)) & 3u);
          const uint32_t bit = 1u << s;
          if (used & bit) {  // wait for the MMA warp's commit that releases the slot's previous occupant
            mbar_wait(&emptyM[s], (phM >> s) & 1);
            phM ^= bit;
          }
          used |= bit;
          unsigned char* st = stage_mem + s * unit_bytes;
          if (k == 0) MUCON_TR16(0, i);
          mbar_arrive_expect_tx(&fullS[s], static_cast<uint32_t>(NKB * R * 128));
          const int row = static_cast<int>(tl.row0) + tl.t0 + (slab ? -dil : (tap - 1) * dil);  // may be negative: TMA zero-fills
          for (int kb = 0; kb < NKB; ++kb) tma_load_2d(st + kb * kb_bytes, &tmX, kb * 64, row, &fullS[s]);
          if (++s == nslot) s = 0;
        }
      }
    }
  } else if (warp == 2) {
    // ================================ fix-up warp =================================
    // rows that lie outside the video are Conv1d's zero padding (the TMA brought the neighbouring video's rows)
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < n_my; ++i) {
      const Tile tl = tiles[blockIdx.x + i * gridDim.x];
      uint32_t code;
      const int nu = tile_units(tl, code);
      for (int k = 0; k < nu; ++k) {
        const int tap = static_cast<int>((code >> (2 * k)) & 3u);
        unsigned char* st = stage_mem + s * unit_bytes;
        mbar_wait(&fullS[s], ph);
        if (lane == 0 && k == 0) MUCON_TR16(1, i);
        if (needs_fix(tl, tap)) {
          // unit row r holds time step t0 + first + r
          const int first = slab ? -dil : (tap - 1) * dil;
          const int lo = -(tl.t0 + first);          // rows below lo are before the video
          const int hi = tl.T - (tl.t0 + first);    // rows from hi on are after it
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) zero_pad_rows(st + kb * kb_bytes, lo, hi, R, lane);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&readyS[s]);
        if (++s == nslot) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = instr_desc_bf16(BM, BN, F16);
    constexpr uint32_t idesc16 = instr_desc_bf16(BM, 16, F16);   // the residual MMAs: N = 16
    const uint32_t w_addr = smem_u32(w_mem);
    const uint64_t b0 = smem_desc(w_addr);
    const uint64_t bid = smem_desc(smem_u32(id_mem));
    if (n_my > 0) mbar_wait(wfull, 0);
    int us = 0;
    uint32_t uph = 0;
    for (int i = 0; i < n_my + LA; ++i) {
      if (i < n_my) {
        // ---- GEMM 1 of tile i: dilated conv, 3 taps x 2 k-blocks x 4 instructions (M128 N128 K16)
        const Tile tl = tiles[blockIdx.x + i * gridDim.x];
        const int acc = i & 1;
        if (lane == 0) MUCON_TR16(12, i);
        mbar_wait(&a1free[acc], ((i >> 1) & 1) ^ 1);
        uint32_t code;
        const int nu = tile_units(tl, code);
        const uint32_t d1 = tmem_base + acc * BN;
        uint32_t issued = 0;
        int cslot = 0;   // slot of the unit that holds the centre tap (the slab, or the tile's last unit)
        for (int ku = 0; ku < nu; ++ku) {
          const int utap = static_cast<int>((code >> (2 * ku)) & 3u);
          if (ku == nu - 1) cslot = us;
          mbar_wait(needs_fix(tl, utap) ? &readyS[us] : &fullS[us], uph);
          tc_fence_after();
          if (lane == 0 && ku == 0) MUCON_TR16(2, i);
          if (lane == 0) {
            // A descriptors: the constant fields plus the start address >> 4.  Slab mode: tap `tap` = rows tap*dil ..
            // +127 of the slab: the SWIZZLE_128B descriptor's start address moved by whole 128-byte rows (the swizzle
            // is a function of the absolute shared-memory address)
            const uint64_t a0 = smem_desc(smem_u32(stage_mem + us * unit_bytes));
            if (slab) {
#pragma unroll
              for (int tap = 0; tap < 3; ++tap) {
                if (!tap_live((tap - 1) * dil, tl.T)) continue;
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                  const uint64_t adesc = a0 + static_cast<uint32_t>((kb * kb_bytes + tap * dil * 128) >> 4);
                  const uint64_t bdesc = b0 + static_cast<uint32_t>(((tap * NKB + kb) * WTILE) >> 4);
#pragma unroll
                  for (int k = 0; k < 4; ++k) mma_bf16(d1, adesc + 2 * k, bdesc + 2 * k, idesc, issued | k);
                  issued = 1;
                }
              }
            } else {
#pragma unroll
              for (int kb = 0; kb < NKB; ++kb) {
                const uint64_t adesc = a0 + static_cast<uint32_t>((kb * kb_bytes) >> 4);
                const uint64_t bdesc = b0 + static_cast<uint32_t>(((utap * NKB + kb) * WTILE) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_bf16(d1, adesc + 2 * k, bdesc + 2 * k, idesc, issued | k);
                issued = 1;
              }
              if (ku != nu - 1) mma_commit(&emptyM[us]);  // a side tap's tile is free as soon as these MMAs retire
            }
          }
          __syncwarp();
          if (++us == nslot) { us = 0; uph ^= 1; }
        }
        if (lane == 0) {
          mma_commit(&a1full[acc]);
          MUCON_TR16(3, i);
        }
        __syncwarp();
        // ---- the residual on the tensor core: acc2 = X_centre . I  (exact: every product is x * 1 or x * 0), as eight
        // M128 N16 K16 instructions against one [16 x 16] identity -- instruction j of a k-block pairs channels 16j ..
        // 16j+15 of the centre rows with the identity's k = 0 .. 15 and writes accumulator columns 16j .. 16j+15; GEMM 2
        // accumulates on top of it, b1 is added by epilogue 2.
        // The tile's centre rows are still in their slot (released by the commit below); accumulator 2 must have been
        // drained by epilogue 2 of tile i - 2.
        mbar_wait(&a2free[acc], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t ac = smem_desc(smem_u32(stage_mem + cslot * unit_bytes)) +
                              static_cast<uint32_t>((slab ? dil * 128 : 0) >> 4);
          const uint32_t d2 = tmem_base + 2 * BN + acc * BN;
          if (!no_res) {
#pragma unroll
            for (int hh = 0; hh < NKB; ++hh) {
              const uint64_t adesc = ac + static_cast<uint32_t>((hh * kb_bytes) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma_bf16(d2 + hh * 64 + 16 * k, adesc + 2 * k, bid, idesc16, 0u);
            }
          }
          mma_commit(&emptyM[cslot]);   // the centre unit / slab may be refilled once these MMAs have retired
        }
        __syncwarp();
      }
      if (i >= LA) {
        // ---- GEMM 2 of tile j: acc2 (= x, from the residual MMAs) += Y . W1^T, Y = relu(acc1 + bd) as bf16 in
        // the first 64 columns of accumulator 1
        const int j = i - LA;
        const int acc = j & 1;
        if (lane == 0) MUCON_TR16(13, j);
        mbar_wait(&yready[acc], (j >> 1) & 1);
        tc_fence_after();
        if (lane == 0) MUCON_TR16(4, j);
        if (lane == 0) {
          const uint32_t d2 = tmem_base + 2 * BN + acc * BN;
          const uint32_t ya = tmem_base + acc * BN;
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // 8 x K16: 8 columns of packed bf16 pairs each
            const uint64_t bdesc = b0 + static_cast<uint32_t>(((3 * NKB + (k >> 2)) * WTILE) >> 4) + 2 * (k & 3);
            mma_bf16_ts(d2, ya + 8 * k, bdesc, idesc, (no_res && k == 0) ? 0u : 1u);
          }
          mma_commit(&a2full[acc]);
          mma_commit(&a1free[acc]);
          MUCON_TR16(5, j);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    EpiCtx c;
    c.tiles = tiles; c.stage_mem = stage_mem; c.staging = staging;
    c.a1full = a1full; c.yready = yready; c.a2full = a2full; c.a2free = a2free;
    c.tmO = &tmO; c.out = out; c.tmem_base = tmem_base;
    c.n_my = n_my; c.nslot = nslot; c.LA = LA; c.unit_bytes = unit_bytes; c.kb_bytes = kb_bytes; c.crow0 = slab ? dil : 0;
    c.slab = slab; c.dil = dil;
    c.S = S; c.pool = pool; c.relu_final = relu_final; c.out_f32 = out_f32;
    if (warp < 8) epilogue_warps<0, F16>(c, bias);
    else epilogue_warps<1, F16>(c, bias);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace layer16
}  // namespace mucon
