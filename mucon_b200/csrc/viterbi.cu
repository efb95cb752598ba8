// viterbi.cu -- transcript-constrained Viterbi alignment for sm_100a.
//
// Replaces reference src/core/viterbi/viterbi.py:49-158 (Viterbi.decode) as called from
// src/mucon/evaluators.py:178-180.  Two kernels:
//
//   scan_*_kernel   per video: sequential cumulative sum over frames per class column in the
//                   input dtype (np.cumsum, viterbi.py:51) and block-score differences
//                   (viterbi.py:68-72).  HBM-bound: reads every log-prob exactly once.  Frame
//                   slabs arrive in shared memory through 1-D TMA bulk copies (UBLKCP) on an
//                   mbarrier ring; one thread per class column walks the slab.
//   dp_kernel       per (video, candidate transcript): the K x N x J dynamic program
//                   (viterbi.py:92-138).  A transcript segment's J hypothesis scores are a SHIFT
//                   REGISTER over a group of 8 lanes x 9 registers (or a whole warp for the long
//                   tail): ageing is a register move + add, the (value, age) arg-max is an in-lane
//                   tree followed by a lane butterfly with the reference's "last writer wins" tie
//                   rule (viterbi.py:26-28); back-pointers, traceback (viterbi.py:140-158) and a
//                   vectorised label writer (viterbi_dp.cuh).  The generic kernel (viterbi_generic.cuh)
//                   keeps the older circular-buffer form in a global-memory workspace for J > 128.
//
// This translation unit is compiled with -fmad=false: every add/sub/mul must round exactly like
// the NumPy scalar arithmetic it replaces.
#include <math.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"
#include "viterbi_dp.cuh"
#include "viterbi_fused.cuh"
#include "viterbi_generic.cuh"
#include "viterbi_lanes.cuh"

namespace mucon {
namespace {

__host__ __device__ __forceinline__ int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

// ============================================================================================
// Block-score scan
// ============================================================================================

// TMA-staged variant.  Requires C*sizeof(T) % 16 == 0 and logp 16-byte aligned.
// smem: [stages] mbarriers, then stages x slab (slab = bps blocks of fs rows of C values).
template <typename T, int FS>
__global__ void __launch_bounds__(128) scan_bulk_kernel(const T* __restrict__ logp, const int64_t* __restrict__ vid_off,
                                                        const int64_t* __restrict__ blk_off,
                                                        const int32_t* __restrict__ order, int C, int fs_rt, int bps,
                                                        int stages, T* __restrict__ bs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int fs = FS ? FS : fs_rt;
  const int v = order ? order[blockIdx.x] : blockIdx.x;
  const int64_t r0 = vid_off[v];
  const int64_t k_base = blk_off[v];
  const int64_t K = blk_off[v + 1] - k_base;
  if (K <= 0) return;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  T* slabs = reinterpret_cast<T*>(smem_raw + 128);
  const uint32_t row_bytes = static_cast<uint32_t>(C) * sizeof(T);
  const uint32_t slab_elems = static_cast<uint32_t>(bps) * fs * C;
  const int64_t nslabs = (K + bps - 1) / bps;
  const unsigned char* src = reinterpret_cast<const unsigned char*>(logp) + r0 * row_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int64_t slab, int st) {
    const int64_t b0 = slab * bps;
    const int nb = static_cast<int>(min64(bps, K - b0));
    const uint32_t bytes = static_cast<uint32_t>(nb) * fs * row_bytes;
    mbar_arrive_expect_tx(&bars[st], bytes);
    bulk_g2s(slabs + static_cast<size_t>(st) * slab_elems, src + b0 * fs * static_cast<int64_t>(row_bytes), bytes,
             &bars[st]);
  };
  if (threadIdx.x == 0) {
    const int pre = static_cast<int>(min64(stages, nslabs));
    for (int s = 0; s < pre; ++s) issue(s, s);
  }

  const int c = threadIdx.x;
  const bool active = c < C;
  T run = neg_zero<T>();  // -0 is the exact additive identity (F[0] = logp[0])
  T prev = 0;
  T* out = bs + k_base * C + c;

  int st = 0;
  uint32_t parity = 0;
  for (int64_t i = 0; i < nslabs; ++i) {
    mbar_wait(&bars[st], parity);
    if (active) {
      const int64_t b0 = i * bps;
      const int nb = static_cast<int>(min64(bps, K - b0));
      const T* s = slabs + static_cast<size_t>(st) * slab_elems + c;
      for (int b = 0; b < nb; ++b) {
        if (FS) {
#pragma unroll
          for (int r = 0; r < (FS ? FS : 1); ++r) run = run + s[r * C];
        } else {
#pragma unroll 4
          for (int r = 0; r < fs; ++r) run = run + s[r * C];
        }
        s += fs * C;
        const T o = (b0 + b == 0) ? run : run - prev;
        prev = run;
        out[(b0 + b) * C] = o;
      }
    }
    __syncthreads();  // everyone is done with stage st
    if (threadIdx.x == 0 && i + stages < nslabs) issue(i + stages, st);
    if (++st == stages) { st = 0; parity ^= 1; }
  }
}

// Direct variant: every thread streams its own class column from global memory (coalesced
// across the row).  Any C, any alignment.
template <typename T>
__global__ void __launch_bounds__(128) scan_direct_kernel(const T* __restrict__ logp,
                                                          const int64_t* __restrict__ vid_off,
                                                          const int64_t* __restrict__ blk_off,
                                                          const int32_t* __restrict__ order, int C, int fs,
                                                          T* __restrict__ bs) {
  const int v = order ? order[blockIdx.x] : blockIdx.x;
  const int64_t r0 = vid_off[v];
  const int64_t k_base = blk_off[v];
  const int64_t K = blk_off[v + 1] - k_base;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const T* p = logp + r0 * C + c;
    T* out = bs + k_base * C + c;
    T run = neg_zero<T>();
    T prev = 0;
    for (int64_t k = 0; k < K; ++k) {
      int r = 0;
      for (; r + 10 <= fs; r += 10) {
        T x[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) x[q] = __ldg(p + static_cast<int64_t>(q) * C);
#pragma unroll
        for (int q = 0; q < 10; ++q) run = run + x[q];
        p += static_cast<int64_t>(10) * C;
      }
      for (; r < fs; ++r) {
        run = run + __ldg(p);
        p += C;
      }
      const T o = (k == 0) ? run : run - prev;
      prev = run;
      out[k * C] = o;
    }
  }
}

// ============================================================================================
// Candidate arg-max and standalone label writer
// ============================================================================================

__global__ void select_kernel(const double* __restrict__ score, const int32_t* __restrict__ status,
                              const int32_t* __restrict__ cand_off, int V, int32_t* __restrict__ best) {
  // one warp per video; lowest index wins ties
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= V) return;
  const int lane = threadIdx.x & 31;
  const int a = cand_off[v], e = cand_off[v + 1];
  double bv = 0.0;
  int bi = -1;
  for (int i = a + lane; i < e; i += 32) {
    if (status[i] == MUCON_UNIT_INFEASIBLE) continue;
    const double x = score[i];
    if (bi < 0 || x > bv) { bv = x; bi = i; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
  }
  if (lane == 0) best[v] = bi;
}

// The same arg-max with the reference's tie order between candidates.  Its hypothesis table is an insertion-ordered
// dict and finalize_decoding (viterbi.py:125-138) takes `score >= best`, so of several final hypotheses with EXACTLY the
// same score the one latest in the dict wins.  The dict order is structural (decode_frame, viterbi.py:93-123: for every
// old key first "stay", then the successors in the grammar's iteration order; an existing key keeps its place): a
// final key (transcript p of d labels, last segment of l blocks) sits at the position of the choice sequence
// [p[0]; stay x (K - l - d + 1); p[1], .., p[d-1]; stay x (l - 1)], compared lexicographically with stay < any
// successor.  Hence: later = larger rank of p[0], then larger l + d, then larger rank sequence of p[1:] (a proper
// prefix is earlier).  tie_rank[2u] = rank of p[0], tie_rank[2u + 1] = dense rank of p[1:] among the video's
// candidates, both built on the host from the grammar's own successor sets (grammar.tie_ranks).
struct SelKey {
  double v;
  int r0, ld, r1, i;
};
__device__ __forceinline__ bool sel_later(const SelKey& a, const SelKey& b) {   // a beats b
  if (b.i < 0) return a.i >= 0;
  if (a.i < 0) return false;
  if (a.v != b.v) return a.v > b.v;
  if (a.r0 != b.r0) return a.r0 > b.r0;
  if (a.ld != b.ld) return a.ld > b.ld;
  if (a.r1 != b.r1) return a.r1 > b.r1;
  return a.i < b.i;
}
__global__ void select_ranked_kernel(const double* __restrict__ score, const int32_t* __restrict__ status,
                                     const int32_t* __restrict__ cand_off, int V, const int32_t* __restrict__ final_j,
                                     const int32_t* __restrict__ tr_off, const int32_t* __restrict__ tie_rank,
                                     int32_t* __restrict__ best) {
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= V) return;
  const int lane = threadIdx.x & 31;
  const int a = cand_off[v], e = cand_off[v + 1];
  SelKey b{0.0, 0, 0, 0, -1};
  for (int i = a + lane; i < e; i += 32) {
    if (status[i] == MUCON_UNIT_INFEASIBLE) continue;
    const SelKey x{score[i], tie_rank[2 * i], final_j[i] + (tr_off[i + 1] - tr_off[i]), tie_rank[2 * i + 1], i};
    if (sel_later(x, b)) b = x;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    SelKey o;
    o.v = __shfl_xor_sync(0xffffffffu, b.v, off);
    o.r0 = __shfl_xor_sync(0xffffffffu, b.r0, off);
    o.ld = __shfl_xor_sync(0xffffffffu, b.ld, off);
    o.r1 = __shfl_xor_sync(0xffffffffu, b.r1, off);
    o.i = __shfl_xor_sync(0xffffffffu, b.i, off);
    if (sel_later(o, b)) b = o;
  }
  if (lane == 0) best[v] = b.i;
}

constexpr int kLabelsMaxN = 128;
__global__ void __launch_bounds__(256) labels_kernel(const int32_t* __restrict__ sel, const int64_t* __restrict__ out_off,
                                                     const int64_t* __restrict__ vid_off,
                                                     const int32_t* __restrict__ unit_vid,
                                                     const int32_t* __restrict__ tr, const int32_t* __restrict__ tr_off,
                                                     const int32_t* __restrict__ seg_blocks, int fs,
                                                     int32_t* __restrict__ labels) {
  __shared__ int trl[kLabelsMaxN];
  __shared__ int64_t segend[kLabelsMaxN];
  __shared__ int last_s;
  const int u = sel[blockIdx.x];
  if (u < 0) return;
  const int v = unit_vid[u];
  const int64_t T = vid_off[v + 1] - vid_off[v];
  const int64_t K = T / fs;
  const int tr0 = tr_off[u], N = tr_off[u + 1] - tr0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) trl[n] = tr[tr0 + n];
  if (threadIdx.x == 0) {
    int64_t pos = T - K * fs;
    int last = 0;
    for (int n = 0; n < N; ++n) {
      const int sb = seg_blocks[tr0 + n];
      if (sb > 0) last = n;
      pos += static_cast<int64_t>(fs) * sb;
      segend[n] = pos;
    }
    last_s = last;
  }
  __syncthreads();
  write_labels(labels + out_off[blockIdx.x], T, T - K * fs, trl, segend, last_s, threadIdx.x, blockDim.x);
}

// ============================================================================================
// Host side
// ============================================================================================

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename T>
int launch_scan(const T* logp, const int64_t* vid_off, const int64_t* blk_off, const int32_t* order, int V, int C,
                int fs, T* bs, cudaStream_t st) {
  const size_t row_bytes = (size_t)C * sizeof(T);
  const bool can_bulk = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(logp) & 15) == 0) && C <= 128;
  const int mode = env_int("MUCON_SCAN_MODE", 0);  // 0 auto, 1 force direct, 2 force bulk
  if (!can_bulk || mode == 1) {
    const int threads = C <= 32 ? 32 : (C <= 64 ? 64 : 128);
    scan_direct_kernel<T><<<V, threads, 0, st>>>(logp, vid_off, blk_off, order, C, fs, bs);
    MUCON_CUDA_CHECK(cudaGetLastError());
    return MUCON_OK;
  }
  const size_t blk_bytes = row_bytes * fs;
  const int slab_target = env_int("MUCON_SCAN_SLAB_BYTES", 23040);
  int bps = (int)(slab_target / blk_bytes);
  if (bps < 1) bps = 1;
  int stages = env_int("MUCON_SCAN_STAGES", 3);
  size_t smem = 128 + (size_t)stages * bps * blk_bytes;
  while (smem > 200 * 1024 && stages > 2) { --stages; smem = 128 + (size_t)stages * bps * blk_bytes; }
  if (smem > 227 * 1024) {  // a single block of frames does not fit: stream from global instead
    const int threads = C <= 32 ? 32 : (C <= 64 ? 64 : 128);
    scan_direct_kernel<T><<<V, threads, 0, st>>>(logp, vid_off, blk_off, order, C, fs, bs);
    MUCON_CUDA_CHECK(cudaGetLastError());
    return MUCON_OK;
  }
  const int threads = C <= 32 ? 32 : (C <= 64 ? 64 : 128);
  auto kern = (fs == 30) ? scan_bulk_kernel<T, 30> : scan_bulk_kernel<T, 0>;
  if (smem > 48 * 1024)
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<V, threads, smem, st>>>(logp, vid_off, blk_off, order, C, fs, bps, stages, bs);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

template <typename BST, int G, int SL>
int launch_dp(const mucon_viterbi_batch& b, int J, cudaStream_t st) {
  // keep the back-pointer table in shared memory when it fits next to everything else
  int bp_rows = b.max_K;
  size_t smem = dp_layout(b.wpc, G, J, sizeof(BST), bp_rows).total;
  if (smem > (size_t)(200 * 1024 / (kDpMaxWarps / b.wpc))) {
    bp_rows = 0;
    smem = dp_layout(b.wpc, G, J, sizeof(BST), 0).total;
  }
  auto kern = dp_kernel<BST, G, SL>;
  if (smem > 48 * 1024)
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<b.n_cta, b.wpc * 32, smem, st>>>(b, J, b.warp_unit, bp_rows);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

template <typename BST>
int dispatch_sl(const mucon_viterbi_batch& b, int J, cudaStream_t st) {
  const int G = b.lanes;
  const int SL = (J + G - 1) / G;
#define MUCON_SL_CASE(g, n) case n: return launch_dp<BST, g, n>(b, J, st);
#ifdef MUCON_ONLY_SL9
  if (G == 8 && SL == 9) return launch_dp<BST, 8, 9>(b, J, st);
  return MUCON_EUNSUPPORTED;
#else
  if (G == 32) {
    switch (SL) {
      MUCON_SL_CASE(32, 1) MUCON_SL_CASE(32, 2) MUCON_SL_CASE(32, 3) MUCON_SL_CASE(32, 4)
      default: return MUCON_EUNSUPPORTED;
    }
  }
  if (G == 4) {
    switch (SL) {
      MUCON_SL_CASE(4, 1) MUCON_SL_CASE(4, 2) MUCON_SL_CASE(4, 3) MUCON_SL_CASE(4, 4)
      MUCON_SL_CASE(4, 5) MUCON_SL_CASE(4, 6) MUCON_SL_CASE(4, 7) MUCON_SL_CASE(4, 8)
      default: return MUCON_EUNSUPPORTED;
    }
  }
  switch (SL) {
    MUCON_SL_CASE(8, 5) MUCON_SL_CASE(8, 6) MUCON_SL_CASE(8, 7) MUCON_SL_CASE(8, 8)
    MUCON_SL_CASE(8, 9) MUCON_SL_CASE(8, 10) MUCON_SL_CASE(8, 11) MUCON_SL_CASE(8, 12)
    MUCON_SL_CASE(8, 13) MUCON_SL_CASE(8, 14) MUCON_SL_CASE(8, 15) MUCON_SL_CASE(8, 16)
    default: return MUCON_EUNSUPPORTED;
  }
#endif
#undef MUCON_SL_CASE
}

template <typename BST, int G, int SL>
int launch_fused(const mucon_viterbi_batch& b, int J, const BST* logp, const int32_t* order, int write_bs,
                 cudaStream_t st, bool pdl, const int64_t* z_off) {
  FusedCfg cfg;
  cfg.z_off = z_off;
  cfg.scan_threads = b.C <= 32 ? 32 : (b.C <= 64 ? 64 : 128);
  const int spw = 32 / G;
  cfg.dp_warps = (b.max_N - 1 + spw - 1) / spw;
  if (cfg.dp_warps < 1) cfg.dp_warps = 1;
  if (cfg.scan_threads + 32 * cfg.dp_warps > (G == 32 ? kFusedMaxThreadsWide : kFusedMaxThreads))
    return MUCON_EUNSUPPORTED;
  const size_t blk_bytes = (size_t)b.C * sizeof(BST) * b.fs;
  cfg.bps = (int)(env_int("MUCON_FUSED_SLAB_BYTES", 17280) / blk_bytes);
  if (cfg.bps < 1) cfg.bps = 1;
  cfg.stages = env_int("MUCON_FUSED_STAGES", 2);
  cfg.ring_slabs = env_int("MUCON_FUSED_RING", 4);
  if (z_off) cfg.stages = 0;  // pooled source: no TMA ring, the scan reads the small table through L1
  if ((!z_off && cfg.stages < 2) || cfg.stages > 8 || cfg.ring_slabs < 2 || cfg.ring_slabs > 8) return MUCON_EINVAL;
  cfg.write_bs = write_bs && b.bs;
  cfg.bp_rows = b.max_K;
  size_t smem = fused_smem_bytes(cfg, G, J, b.C, b.fs, sizeof(BST));
  if (smem > 100 * 1024) {  // a long back-pointer table would cost residency: trace from HBM instead
    cfg.bp_rows = 0;
    smem = fused_smem_bytes(cfg, G, J, b.C, b.fs, sizeof(BST));
  }
  if (smem > 227 * 1024) return MUCON_EUNSUPPORTED;
  void (*kern)(const mucon_viterbi_batch, const int, const BST*, const int32_t*, const FusedCfg) =
      (b.fs == 30) ? align_fused_kernel<BST, G, SL, 30> : align_fused_kernel<BST, G, SL, 0>;
  if constexpr (sizeof(BST) == 4 && G == 8 && SL == 9) {
    // the evaluator's shape (float32, fs = 30, J = 66) with 33..64 classes: one scan warp, two
    // class columns per lane
    if (b.fs == 30 && b.C > 32 && b.C <= 64 && b.C % 2 == 0 && env_int("MUCON_FUSED_CPT", 2) == 2) {
      cfg.scan_threads = 32;
      kern = align_fused_kernel<BST, G, SL, 30, kFusedMaxThreads, 2, 2>;  // <= 128 registers: 4 CTAs per SM
      if (b.C == 48) kern = align_fused_kernel<BST, G, SL, 30, kFusedMaxThreads, 2, 2, 48>;  // Breakfast
    }
  }
  if (smem > 48 * 1024)
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (pdl) {
    // programmatic dependent launch: this grid may start as soon as every CTA of the previous kernel
    // in the stream has executed griddepcontrol.launch_dependents (align_fused_kernel does so at its
    // start), i.e. it runs CONCURRENTLY with that kernel.  There is no data dependency between the
    // two (disjoint units), so this kernel never executes griddepcontrol.wait.
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(b.U);
    lc.blockDim = dim3(cfg.scan_threads + 32 * cfg.dp_warps);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    MUCON_CUDA_CHECK(cudaLaunchKernelEx(&lc, kern, b, J, logp, order, cfg));
  } else {
    kern<<<b.U, cfg.scan_threads + 32 * cfg.dp_warps, smem, st>>>(b, J, logp, order, cfg);
  }
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

template <typename BST>
int dispatch_fused(const mucon_viterbi_batch& b, int J, const BST* logp, const int32_t* order, int write_bs,
                   cudaStream_t st, bool pdl, const int64_t* z_off) {
  const int G = b.lanes == 32 ? 32 : (J <= 32 ? 4 : 8);
  const int SL = (J + G - 1) / G;
#define MUCON_SL_CASE(g, n) case n: return launch_fused<BST, g, n>(b, J, logp, order, write_bs, st, pdl, z_off);
#ifdef MUCON_ONLY_SL9  // developer builds: only the evaluator's shape (J = 66)
  if (G == 8 && SL == 9) return launch_fused<BST, 8, 9>(b, J, logp, order, write_bs, st, pdl, z_off);
  return MUCON_EUNSUPPORTED;
#else
  if (G == 32) {
    switch (SL) {
      MUCON_SL_CASE(32, 1) MUCON_SL_CASE(32, 2) MUCON_SL_CASE(32, 3) MUCON_SL_CASE(32, 4)
      default: return MUCON_EUNSUPPORTED;
    }
  }
  if (G == 4) {
    switch (SL) {
      MUCON_SL_CASE(4, 1) MUCON_SL_CASE(4, 2) MUCON_SL_CASE(4, 3) MUCON_SL_CASE(4, 4)
      MUCON_SL_CASE(4, 5) MUCON_SL_CASE(4, 6) MUCON_SL_CASE(4, 7) MUCON_SL_CASE(4, 8)
      default: return MUCON_EUNSUPPORTED;
    }
  }
  switch (SL) {
    MUCON_SL_CASE(8, 5) MUCON_SL_CASE(8, 6) MUCON_SL_CASE(8, 7) MUCON_SL_CASE(8, 8)
    MUCON_SL_CASE(8, 9) MUCON_SL_CASE(8, 10) MUCON_SL_CASE(8, 11) MUCON_SL_CASE(8, 12)
    MUCON_SL_CASE(8, 13) MUCON_SL_CASE(8, 14) MUCON_SL_CASE(8, 15) MUCON_SL_CASE(8, 16)
    default: return MUCON_EUNSUPPORTED;
  }
#endif
#undef MUCON_SL_CASE
}

int warps_for(int N, int G) {
  const int spw = 32 / G;
  const int w = (N - 1 + spw - 1) / spw;
  return w < 1 ? 1 : w;
}

}  // namespace
}  // namespace mucon

using namespace mucon;

extern "C" int mucon_viterbi_blockscores(const void* logp, int in_is_f64, const int64_t* vid_off,
                                         const int64_t* blk_off, const int32_t* order, int V, int C, int fs,
                                         void* bs, void* stream) {
  if (!logp || !vid_off || !blk_off || !bs || V < 0 || C < 1 || fs < 1) return MUCON_EINVAL;
  if (V == 0) return MUCON_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_is_f64)
    return launch_scan<double>(static_cast<const double*>(logp), vid_off, blk_off, order, V, C, fs,
                               static_cast<double*>(bs), st);
  return launch_scan<float>(static_cast<const float*>(logp), vid_off, blk_off, order, V, C, fs,
                            static_cast<float*>(bs), st);
}

extern "C" int mucon_viterbi_decode(const mucon_viterbi_batch* bh, void* stream) {
  if (!bh) return MUCON_EINVAL;
  const mucon_viterbi_batch& b = *bh;
  if (b.U < 0 || b.C < 1 || b.fs < 1 || b.max_len < b.fs || b.max_N < 1 || b.n_cta < 0 || b.max_K < 0)
    return MUCON_EINVAL;
  if (!b.bs || !b.vid_off || !b.blk_off || !b.unit_vid || !b.tr || !b.tr_off || !b.score || !b.seg_blocks ||
      !b.final_j || !b.status || !b.bp || !b.bp_off || !b.warp_unit)
    return MUCON_EINVAL;
  if (!b.len_rows && !(b.len_params && b.logfact)) return MUCON_EINVAL;
  if (b.U == 0 || b.n_cta == 0) return MUCON_OK;
  const int J = b.max_len / b.fs;
  if (J > kDpMaxJ) return MUCON_EUNSUPPORTED;  // ages live in registers, back-pointers are uint8
  if (b.lanes != dp_group(J, b.max_N, b.lanes)) return MUCON_EINVAL;
  if (b.max_N > dp_max_n(b.lanes)) return MUCON_EUNSUPPORTED;
  if (b.wpc != 4 && b.wpc != 8 && b.wpc != 16) return MUCON_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (b.bs_is_f64) return dispatch_sl<double>(b, J, st);
  return dispatch_sl<float>(b, J, st);
}

extern "C" int mucon_viterbi_decode_generic(const mucon_viterbi_batch* bh, double* ws, const int64_t* ws_off,
                                            int bp_is_u16, void* stream) {
  if (!bh || !ws || !ws_off) return MUCON_EINVAL;
  const mucon_viterbi_batch& b = *bh;
  if (b.U < 0 || b.C < 1 || b.fs < 1 || b.max_len < b.fs || b.max_N < 1) return MUCON_EINVAL;
  if (!b.bs || !b.vid_off || !b.blk_off || !b.unit_vid || !b.tr || !b.tr_off || !b.score || !b.seg_blocks ||
      !b.final_j || !b.status || !b.bp || !b.bp_off)
    return MUCON_EINVAL;
  if (!b.len_rows && !(b.len_params && b.logfact)) return MUCON_EINVAL;
  if (b.U == 0) return MUCON_OK;
  const int J = b.max_len / b.fs;
  if (!bp_is_u16 && J > 255) return MUCON_EINVAL;
  if (J > 65535) return MUCON_EUNSUPPORTED;
  const size_t smem = (size_t)b.max_N * (sizeof(double) + sizeof(int64_t) + 3 * sizeof(int));
  if (smem > 200 * 1024) return MUCON_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MUCON_GEN(BST, BPT)                                                                                   \
  do {                                                                                                        \
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(dp_generic_kernel<BST, BPT>,                                        \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    dp_generic_kernel<BST, BPT><<<b.U, kGenThreads, smem, st>>>(b, J, ws, ws_off);                            \
  } while (0)
  if (b.bs_is_f64) { if (bp_is_u16) MUCON_GEN(double, uint16_t); else MUCON_GEN(double, uint8_t); }
  else { if (bp_is_u16) MUCON_GEN(float, uint16_t); else MUCON_GEN(float, uint8_t); }
#undef MUCON_GEN
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_viterbi_pack_lanes_h(const int32_t* N_h, const int32_t* order_h, int U, int32_t* lane_unit_h,
                                          int32_t* n_warps_out) {
  if (!N_h || !lane_unit_h || !n_warps_out || U < 0) return MUCON_EINVAL;
  // first fit over a window of open warps, units taken in order_h (longest first), so that the
  // units sharing a warp have similar numbers of steps
  constexpr int kWindow = 4;
  int open_warp[kWindow], open_used[kWindow], n_open = 0, n_warps = 0;
  for (int i = 0; i < U; ++i) {
    const int u = order_h ? order_h[i] : i;
    const int need = N_h[u] > 1 ? N_h[u] - 1 : 1;
    if (need > 32) return MUCON_EUNSUPPORTED;
    int pick = -1;
    for (int o = 0; o < n_open; ++o)
      if (32 - open_used[o] >= need) { pick = o; break; }
    if (pick < 0) {
      if (n_open == kWindow) {
        for (int o = 1; o < kWindow; ++o) { open_warp[o - 1] = open_warp[o]; open_used[o - 1] = open_used[o]; }
        --n_open;
      }
      pick = n_open++;
      open_warp[pick] = n_warps++;
      open_used[pick] = 0;
      for (int l = 0; l < 32; ++l) lane_unit_h[(size_t)open_warp[pick] * 32 + l] = -1;
    }
    for (int l = 0; l < need; ++l) lane_unit_h[(size_t)open_warp[pick] * 32 + open_used[pick] + l] = u;
    open_used[pick] += need;
  }
  *n_warps_out = n_warps;
  return MUCON_OK;
}

extern "C" int mucon_viterbi_decode_lanes(const mucon_viterbi_batch* bh, const int32_t* lane_unit, int n_warps,
                                          const int32_t* reserved, void* stream) {
  if (!bh || !lane_unit || n_warps < 0 || reserved) return MUCON_EINVAL;
  const int32_t* progress = nullptr;
  const mucon_viterbi_batch& b = *bh;
  if (b.U < 0 || b.C < 1 || b.fs < 1 || b.max_len < b.fs || b.max_N < 1) return MUCON_EINVAL;
  if (!b.bs || !b.vid_off || !b.blk_off || !b.unit_vid || !b.tr || !b.tr_off || !b.score || !b.seg_blocks ||
      !b.final_j || !b.status || !b.bp || !b.bp_off)
    return MUCON_EINVAL;
  if (!b.len_rows && !(b.len_params && b.logfact)) return MUCON_EINVAL;
  if (b.U == 0 || n_warps == 0) return MUCON_OK;
  const int J = b.max_len / b.fs;
  if (J > 66 || b.max_N > 33) return MUCON_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MUCON_LANES(BST)                                                                                     \
  do {                                                                                                       \
    const size_t smem = lanes_layout(66, sizeof(BST)).total;                                                 \
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(dp_lanes_kernel<BST, 6, 11>,                                       \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    dp_lanes_kernel<BST, 6, 11><<<n_warps, 32, smem, st>>>(b, J, lane_unit, progress);                       \
  } while (0)
  if (b.bs_is_f64) MUCON_LANES(double); else MUCON_LANES(float);
#undef MUCON_LANES
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

static int align_fused_impl(const mucon_viterbi_batch* bh, const void* logp, int in_is_f64, const int32_t* order,
                            int write_bs, void* stream, bool pdl, const int64_t* z_off = nullptr) {
  if (!bh || !logp) return MUCON_EINVAL;
  const mucon_viterbi_batch& b = *bh;
  if (b.U < 0 || b.C < 1 || b.fs < 1 || b.max_len < b.fs || b.max_N < 1 || b.max_K < 0) return MUCON_EINVAL;
  if (!b.vid_off || !b.blk_off || !b.unit_vid || !b.tr || !b.tr_off || !b.score || !b.seg_blocks || !b.final_j ||
      !b.status || !b.bp || !b.bp_off)
    return MUCON_EINVAL;
  if (!b.len_rows && !(b.len_params && b.logfact)) return MUCON_EINVAL;
  if ((in_is_f64 != 0) != (b.bs_is_f64 != 0)) return MUCON_EINVAL;
  if (b.U == 0) return MUCON_OK;
  const int J = b.max_len / b.fs;
  if (J > kDpMaxJ) return MUCON_EUNSUPPORTED;
  const size_t row_bytes = (size_t)b.C * (in_is_f64 ? 8 : 4);
  if (row_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(logp) & 15) != 0 || b.C > 128) return MUCON_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_is_f64) return dispatch_fused<double>(b, J, static_cast<const double*>(logp), order, write_bs, st, pdl, z_off);
  return dispatch_fused<float>(b, J, static_cast<const float*>(logp), order, write_bs, st, pdl, z_off);
}

extern "C" int mucon_viterbi_align_fused(const mucon_viterbi_batch* bh, const void* logp, int in_is_f64,
                                         const int32_t* order, int write_bs, void* stream) {
  return align_fused_impl(bh, logp, in_is_f64, order, write_bs, stream, false);
}

static int align_fused_tail_impl(const mucon_viterbi_batch* bh, const void* logp, int in_is_f64, const int32_t* order,
                                 int n_wide, int write_bs, void* stream, const int64_t* z_off) {
  if (!bh || !order || n_wide < 0 || n_wide > bh->U) return MUCON_EINVAL;
  mucon_viterbi_batch b = *bh;
  int rc = MUCON_OK;
  if (n_wide > 0) {
    // the wide launch (a warp per transcript segment) for the first n_wide units of `order`
    b.U = n_wide;
    b.lanes = 32;
    rc = align_fused_impl(&b, logp, in_is_f64, order, write_bs, stream, false, z_off);
    if (rc == MUCON_EUNSUPPORTED) { n_wide = 0; rc = MUCON_OK; }  // shape not covered: everything in the main launch
    if (rc != MUCON_OK) return rc;
  }
  b.U = bh->U - n_wide;
  b.lanes = 0;
  // the main launch follows in the same stream as a programmatic dependent launch: it starts once
  // the wide CTAs are resident and runs concurrently with them
  if (b.U > 0) rc = align_fused_impl(&b, logp, in_is_f64, order + n_wide, write_bs, stream, n_wide > 0, z_off);
  return rc;
}

extern "C" int mucon_viterbi_align_fused_tail(const mucon_viterbi_batch* bh, const void* logp, int in_is_f64,
                                              const int32_t* order, int n_wide, int write_bs, void* stream) {
  return align_fused_tail_impl(bh, logp, in_is_f64, order, n_wide, write_bs, stream, nullptr);
}

extern "C" int mucon_viterbi_align_fused_pooled(const mucon_viterbi_batch* bh, const void* logp_z, int in_is_f64,
                                                const int64_t* z_off, const int32_t* order, int n_wide,
                                                int write_bs, void* stream) {
  if (!z_off) return MUCON_EINVAL;
  return align_fused_tail_impl(bh, logp_z, in_is_f64, order, n_wide, write_bs, stream, z_off);
}

// ---------------------------------------------------------------------------------------------
// Single-video session: the reference's own call pattern (evaluators.py:147-180 decodes ONE video per call, batch
// size 1) without per-call allocations or Python-side packing.  One host->device copy carries [metadata | log-
// probabilities] from a pinned staging buffer, the fused kernel runs, one device->host copy brings back
// [score | final_j | status | segment lengths | labels].  Not thread-safe per session.
struct mucon_single {
  int max_T, C, max_N, elem;
  size_t off_tr, off_params, off_logfact, off_logp, stage_bytes;
  size_t out_seg, out_labels, out_bytes;
  unsigned char *h_stage, *d_stage, *h_out, *d_out;
  uint8_t* d_bp;
  int lf_fs, lf_max_len;
  double logfact[kDpMaxJ + 1];
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" int mucon_single_destroy(mucon_single* s) {
  if (!s) return MUCON_EINVAL;
  if (s->h_stage) cudaFreeHost(s->h_stage);
  if (s->h_out) cudaFreeHost(s->h_out);
  if (s->d_stage) cudaFree(s->d_stage);
  if (s->d_out) cudaFree(s->d_out);
  if (s->d_bp) cudaFree(s->d_bp);
  delete s;
  return MUCON_OK;
}

extern "C" int mucon_single_create(int max_T, int C, int max_N, int elem_bytes, mucon_single** out) {
  if (!out || max_T < 1 || C < 1 || max_N < 1 || (elem_bytes != 4 && elem_bytes != 8)) return MUCON_EINVAL;
  mucon_single* s = new mucon_single();
  s->max_T = max_T; s->C = C; s->max_N = max_N; s->elem = elem_bytes;
  s->lf_fs = s->lf_max_len = -1;
  // metadata: vid_off[2] blk_off[2] bp_off[1] lab_off[1] (int64) | unit_vid[1] order[1] tr_off[2] (int32) | tr | params | logfact
  s->off_tr = 64;
  s->off_params = align_up(s->off_tr + 4 * (size_t)max_N, 16);
  s->off_logfact = s->off_params + 24 * (size_t)max_N;
  s->off_logp = align_up(s->off_logfact + 8 * (kDpMaxJ + 1), 128);
  s->stage_bytes = s->off_logp + (size_t)max_T * C * elem_bytes;
  s->out_seg = 16;
  s->out_labels = align_up(s->out_seg + 4 * (size_t)max_N, 16);
  s->out_bytes = s->out_labels + 4 * (size_t)max_T;
  bool ok = cudaMallocHost(reinterpret_cast<void**>(&s->h_stage), s->stage_bytes) == cudaSuccess &&
            cudaMallocHost(reinterpret_cast<void**>(&s->h_out), s->out_bytes) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void**>(&s->d_stage), s->stage_bytes) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void**>(&s->d_out), s->out_bytes) == cudaSuccess &&
            cudaMalloc(reinterpret_cast<void**>(&s->d_bp), (size_t)max_T * max_N + 16) == cudaSuccess;
  if (!ok) {
    mucon::set_cuda_error(cudaGetLastError(), "mucon_single_create");
    mucon_single_destroy(s);
    return MUCON_ECUDA;
  }
  *out = s;
  return MUCON_OK;
}

extern "C" int mucon_single_decode_h(mucon_single* s, const void* logp_h, int is_f64, int T, const int32_t* tr_h, int N,
                                     const double* len_params_h, int fs, int max_len, int seg0_f32, double* score_h,
                                     int32_t* labels_h, int32_t* seg_blocks_h, int32_t* status_h, int32_t* final_j_h,
                                     void* stream) {
  if (!s || !logp_h || !tr_h || !len_params_h || !score_h || !labels_h || !seg_blocks_h || !status_h || !final_j_h ||
      T < 1 || N < 1 || fs < 1 || max_len < fs)
    return MUCON_EINVAL;
  if (T > s->max_T || N > s->max_N || (is_f64 ? 8 : 4) != s->elem) return MUCON_ESHAPE;
  const int J = max_len / fs;
  if (J > kDpMaxJ) return MUCON_EUNSUPPORTED;
  if (s->lf_fs != fs || s->lf_max_len != max_len) {
    int rc = mucon_logfact_h(fs, max_len, s->logfact);
    if (rc != MUCON_OK) return rc;
    s->lf_fs = fs;
    s->lf_max_len = max_len;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int K = T / fs;
  int64_t* m64 = reinterpret_cast<int64_t*>(s->h_stage);
  m64[0] = 0; m64[1] = T;          // vid_off
  m64[2] = 0; m64[3] = K;          // blk_off
  m64[4] = 0;                      // bp_off
  m64[5] = 0;                      // lab_off
  int32_t* m32 = reinterpret_cast<int32_t*>(s->h_stage + 48);
  m32[0] = 0;                      // unit_vid
  m32[1] = 0;                      // order
  m32[2] = 0; m32[3] = N;          // tr_off
  memcpy(s->h_stage + s->off_tr, tr_h, 4 * (size_t)N);
  memcpy(s->h_stage + s->off_params, len_params_h, 24 * (size_t)N);
  memcpy(s->h_stage + s->off_logfact, s->logfact, 8 * (size_t)(J + 1));
  const size_t lp_bytes = (size_t)T * s->C * s->elem;
  memcpy(s->h_stage + s->off_logp, logp_h, lp_bytes);
  // Short videos: the kernel reads the log-probabilities straight from the pinned staging buffer and writes its results
  // into pinned host memory itself (both are mapped into the device's address space under UVA): the 384 KB of a
  // Breakfast video cross PCIe inside the kernel's own TMA ring, overlapped with the DP, and there is no result copy
  // (c1 decode 128 -> 119 us).  MUCON_SINGLE_ZEROCOPY=0 restores the copies, =2 also reads the metadata in place.
  static const int zc_mode = getenv("MUCON_SINGLE_ZEROCOPY") ? atoi(getenv("MUCON_SINGLE_ZEROCOPY")) : 1;
  // 0: copy everything; 1: log-probabilities and results zero-copy, metadata copied; 2: nothing copied at all
  const bool zero_copy = zc_mode != 0 && lp_bytes <= (1u << 20);   // long videos: the copy engine beats the kernel's shallow ring
  const bool meta_zc = zero_copy && zc_mode == 2;
  if (!meta_zc)
    MUCON_CUDA_CHECK(cudaMemcpyAsync(s->d_stage, s->h_stage, zero_copy ? s->off_logp : s->off_logp + lp_bytes,
                                     cudaMemcpyHostToDevice, st));
  mucon_viterbi_batch b;
  memset(&b, 0, sizeof(b));
  b.U = 1; b.C = s->C; b.fs = fs; b.max_len = max_len; b.bs_is_f64 = is_f64 ? 1 : 0; b.seg0_f32 = seg0_f32 ? 1 : 0;
  b.max_N = N; b.max_K = K; b.n_cta = 0; b.wpc = 4; b.lanes = 0;
  unsigned char* d = meta_zc ? s->h_stage : s->d_stage;
  b.vid_off = reinterpret_cast<const int64_t*>(d);
  b.blk_off = reinterpret_cast<const int64_t*>(d + 16);
  b.bp_off = reinterpret_cast<const int64_t*>(d + 32);
  b.lab_off = reinterpret_cast<const int64_t*>(d + 40);
  b.unit_vid = reinterpret_cast<const int32_t*>(d + 48);
  const int32_t* order = reinterpret_cast<const int32_t*>(d + 52);
  b.tr_off = reinterpret_cast<const int32_t*>(d + 56);
  b.tr = reinterpret_cast<const int32_t*>(d + s->off_tr);
  b.len_params = reinterpret_cast<const double*>(d + s->off_params);
  b.logfact = reinterpret_cast<const double*>(d + s->off_logfact);
  unsigned char* o = zero_copy ? s->h_out : s->d_out;
  b.score = reinterpret_cast<double*>(o);
  b.final_j = reinterpret_cast<int32_t*>(o + 8);
  b.status = reinterpret_cast<int32_t*>(o + 12);
  b.seg_blocks = reinterpret_cast<int32_t*>(o + s->out_seg);
  b.labels = reinterpret_cast<int32_t*>(o + s->out_labels);
  b.bp = s->d_bp;
  int rc = align_fused_impl(&b, zero_copy ? s->h_stage + s->off_logp : s->d_stage + s->off_logp, is_f64, order, 0, stream, false);
  if (rc != MUCON_OK) return rc;
  if (!zero_copy)
    MUCON_CUDA_CHECK(cudaMemcpyAsync(s->h_out, s->d_out, s->out_labels + 4 * (size_t)T, cudaMemcpyDeviceToHost, st));
  MUCON_CUDA_CHECK(cudaStreamSynchronize(st));
  memcpy(score_h, s->h_out, 8);
  memcpy(final_j_h, s->h_out + 8, 4);
  memcpy(status_h, s->h_out + 12, 4);
  memcpy(seg_blocks_h, s->h_out + s->out_seg, 4 * (size_t)N);
  memcpy(labels_h, s->h_out + s->out_labels, 4 * (size_t)T);
  return MUCON_OK;
}

// ---------------------------------------------------------------------------------------------
// Evaluator glue on the device (reference src/mucon/evaluators.py:155-165 + core/viterbi/length_model.py:54-63):
// class-mean lengths from the s-head's relative lengths and the Poisson parameters (ln m, m, norms) of every
// transcript position, one warp per video.  lengths[c] = n_frames * sum_{m: tr[m] = c} rel[m] / count(c), exact
// zeros -> 1 (classes that do not occur never reach the DP: only the transcript's positions are produced).
// Sums run over the positions in ascending order in float64, like the float32 . float64 one-hot product they
// replace.  ln() is CUDA's (<= 1 ulp from numpy's): path scores may differ from the host-built parameters in the
// last bits; mucon_b200/evaluate.py keeps the host path as the bit-exact default.
__global__ void __launch_bounds__(128) class_mean_params_kernel(const float* __restrict__ rel, const int32_t* __restrict__ tr,
                                                                const int32_t* __restrict__ tr_off,
                                                                const int64_t* __restrict__ vid_off, int V,
                                                                const double* __restrict__ logtail, int logtail_n,
                                                                double* __restrict__ out) {
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (v >= V) return;
  const int t0 = tr_off[v], N = tr_off[v + 1] - t0;
  const double frames = static_cast<double>(vid_off[v + 1] - vid_off[v]);
  for (int n = lane; n < N; n += 32) {
    const int c = tr[t0 + n];
    double s = 0.0;
    int k = 0;
    for (int m = 0; m < N; ++m)
      if (tr[t0 + m] == c) { s += static_cast<double>(rel[t0 + m]); ++k; }
    double len = s * frames;
    len = len / static_cast<double>(k);
    if (len == 0.0) len = 1.0;
    const double r = rint(len);
    long long mi = static_cast<long long>(len);
    if (mi < 0) mi = 0;
    if (mi >= logtail_n) mi = logtail_n - 1;
    double* o = out + static_cast<size_t>(t0 + n) * 3;
    o[0] = log(len);
    o[1] = len;
    o[2] = (r * log(r) - r) - logtail[mi];
  }
}

extern "C" int mucon_class_mean_params(const float* rel, const int32_t* tr, const int32_t* tr_off, const int64_t* vid_off,
                                       int V, const double* logtail, int logtail_n, double* len_params_out, void* stream) {
  if (!rel || !tr || !tr_off || !vid_off || !logtail || !len_params_out || V < 0 || logtail_n < 2) return MUCON_EINVAL;
  if (V == 0) return MUCON_OK;
  class_mean_params_kernel<<<(V + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(rel, tr, tr_off, vid_off, V, logtail,
                                                                                      logtail_n, len_params_out);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_viterbi_pack_h(const int32_t* N_h, const int32_t* order_h, int U, int max_N, int fs,
                                    int max_len, int lanes, int32_t* warp_unit_h, int32_t* n_cta_out,
                                    int32_t* wpc_out, int32_t* lanes_out) {
  if (!N_h || !warp_unit_h || !n_cta_out || !wpc_out || !lanes_out || U < 0 || max_N < 1 || fs < 1 ||
      max_len < fs)
    return MUCON_EINVAL;
  const int J = max_len / fs;
  if (J > kDpMaxJ) return MUCON_EUNSUPPORTED;
  const int G = dp_group(J, max_N, lanes);
  if (max_N > dp_max_n(G)) return MUCON_EUNSUPPORTED;
  *lanes_out = G;
  // small CTAs keep the grid fine-grained (the block scheduler balances the SMs); a CTA only has
  // to be as large as the largest unit
  const int wmax = warps_for(max_N, G);
  const int wpc = wmax <= 4 ? 4 : (wmax <= 8 ? 8 : 16);
  *wpc_out = wpc;
  constexpr int kWindow = 8;  // open bins that may still take units (keeps K similar within a bin)
  int open_bin[kWindow], open_free[kWindow], n_open = 0;
  int n_cta = 0;
  for (int i = 0; i < U; ++i) {
    const int u = order_h ? order_h[i] : i;
    const int need = warps_for(N_h[u], G);
    int pick = -1;
    for (int o = 0; o < n_open; ++o)
      if (open_free[o] >= need) { pick = o; break; }
    if (pick < 0) {
      if (n_open == kWindow) {  // retire the oldest open bin
        for (int o = 1; o < kWindow; ++o) { open_bin[o - 1] = open_bin[o]; open_free[o - 1] = open_free[o]; }
        --n_open;
      }
      pick = n_open++;
      open_bin[pick] = n_cta++;
      open_free[pick] = wpc;
      for (int w = 0; w < wpc; ++w) warp_unit_h[(size_t)open_bin[pick] * wpc + w] = -1;
    }
    const int start = wpc - open_free[pick];
    for (int w = 0; w < need; ++w) warp_unit_h[(size_t)open_bin[pick] * wpc + start + w] = u;
    open_free[pick] -= need;
  }
  *n_cta_out = n_cta;
  return MUCON_OK;
}

extern "C" int mucon_viterbi_select(const double* score, const int32_t* status, const int32_t* cand_off, int V,
                                    int32_t* best, void* stream) {
  if (!score || !status || !cand_off || !best || V < 0) return MUCON_EINVAL;
  if (V == 0) return MUCON_OK;
  const int wpb = 4;
  select_kernel<<<(V + wpb - 1) / wpb, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(score, status, cand_off, V,
                                                                                          best);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_viterbi_select_ranked(const double* score, const int32_t* status, const int32_t* cand_off, int V,
                                           const int32_t* final_j, const int32_t* tr_off, const int32_t* tie_rank,
                                           int32_t* best, void* stream) {
  if (!score || !status || !cand_off || !final_j || !tr_off || !tie_rank || !best || V < 0) return MUCON_EINVAL;
  if (V == 0) return MUCON_OK;
  const int wpb = 4;
  select_ranked_kernel<<<(V + wpb - 1) / wpb, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      score, status, cand_off, V, final_j, tr_off, tie_rank, best);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_viterbi_labels(const int32_t* sel, int n_sel, const int64_t* out_off, const int64_t* vid_off,
                                    const int32_t* unit_vid, const int32_t* tr, const int32_t* tr_off,
                                    const int32_t* seg_blocks, int fs, int32_t* labels, void* stream) {
  if (!sel || !out_off || !vid_off || !unit_vid || !tr || !tr_off || !seg_blocks || !labels || n_sel < 0 || fs < 1)
    return MUCON_EINVAL;
  if (n_sel == 0) return MUCON_OK;
  labels_kernel<<<n_sel, 256, 0, static_cast<cudaStream_t>(stream)>>>(sel, out_off, vid_off, unit_vid, tr, tr_off,
                                                                       seg_blocks, fs, labels);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_poisson_params_h(const double* means_h, int n, double* out_h) {
  if (!means_h || !out_h || n < 0) return MUCON_EINVAL;
  for (int c = 0; c < n; ++c) {
    const double m = means_h[c];
    const double r = nearbyint(m);  // round-half-even, like np.round
    double norms = r * log(r) - r;
    double lf = 0.0;
    for (long k = 2; k <= (long)m; ++k) lf += log((double)k);
    out_h[c * 3 + 0] = log(m);
    out_h[c * 3 + 1] = m;
    out_h[c * 3 + 2] = norms - lf;
  }
  return MUCON_OK;
}

extern "C" int mucon_logfact_h(int fs, int max_len, double* out_h) {
  if (!out_h || fs < 1 || max_len < fs) return MUCON_EINVAL;
  const int J = max_len / fs;
  double lf = 0.0;
  out_h[0] = 0.0;
  int j = 1;
  for (int l = 1; l <= J * fs; ++l) {
    lf += log((double)l);
    if (l == j * fs) out_h[j++] = lf;
  }
  return MUCON_OK;
}
