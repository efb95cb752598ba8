// viterbi_fused.cuh -- block-score scan and dynamic program of one video in ONE CTA (sm_100a).
//
// For batches with a single transcript per video (the evaluator's case, reference
// src/mucon/evaluators.py:147-180) the scan and the DP are warp-specialised roles of the same
// CTA: the scan warps stream the video's log-probabilities through a TMA ring (UBLKCP +
// mbarrier), keep the sequential per-class running sum (np.cumsum, viterbi.py:51) and publish
// the block scores (viterbi.py:68-72) of every class into a small shared ring; the DP warps
// (dp_unit, viterbi_dp.cuh) consume that ring step by step.  Block scores never travel through
// HBM, the two roles overlap (the scan is HBM-latency-bound, the DP is issue-bound), and the
// whole alignment of a batch is a single launch.
#pragma once
#include "viterbi_dp.cuh"

namespace mucon {

struct FusedCfg {
  int scan_threads;  // 32 / 64 / 128, >= C
  int dp_warps;      // DP warps per CTA (>= warps of the largest unit)
  int bps;           // blocks per slab (TMA slab == ring slab)
  int stages;        // TMA ring depth
  int ring_slabs;    // block-score ring depth
  int bp_rows;       // rows of the shared back-pointer stage (0: trace from HBM)
  int write_bs;      // also store the block scores to b.bs
  // Pooled source (NULL = `logp` holds one row per frame).  Otherwise `logp` holds the log-probabilities at the
  // backbone's pooled resolution ([sum Tz, C], z_off[V+1] row offsets) and frame t of a video reads row
  // min(floor(t * (float)Tz / T), Tz - 1) -- F.interpolate(mode="nearest"), reference src/mucon/models.py:574-576:
  // the scan walks the same float32 sequence as over the expanded [T, C] array, which is never materialised.
  const int64_t* z_off;
};

constexpr int kFusedMaxThreads = 256;      // segments sharing a warp (G = 4 / 8)
constexpr int kFusedMaxThreadsWide = 512;  // a warp per segment (G = 32)
constexpr int kFusedBarBytes = 256;  // mbarriers: tma_full[<=8], ring_full[<=8], ring_empty[<=8]

__host__ __device__ inline size_t fused_smem_bytes(const FusedCfg& c, int G, int J, int C, int fs, int elem) {
  size_t o = kFusedBarBytes;
  o += (size_t)c.stages * c.bps * fs * C * elem;
  o += (size_t)c.ring_slabs * c.bps * C * elem;
  o = (o + 15) & ~size_t(15);
  o += dp_layout(c.dp_warps, G, J, 0, c.bp_rows).total;
  return o;
}

// MAXT / MINB: launch bounds.  The generic instantiations allow 256 (512 for G = 32) threads and
// one CTA per SM; the Breakfast shape (float32, J = 66, fs = 30, <= 13 segments: 160 threads) has a
// dedicated instantiation whose bounds let the compiler target more resident CTAs.
// CPT: class columns per scan thread.  1 = one thread per class (any C <= 128); 2 = one scan WARP
// walks all columns of a 33..64-class problem, two adjacent classes per lane (two independent
// running sums per thread, 8-byte shared loads): one warp less per CTA, so four CTAs fit the
// register file of an SM instead of three.
// CT: compile-time class count (0 = runtime): the scan's 30 shared loads per block then use immediate
// offsets instead of a chain of address additions.
template <typename BST, int G, int SL, int FS, int MAXT = (G == 32 ? kFusedMaxThreadsWide : kFusedMaxThreads), int MINB = 1,
          int CPT = 1, int CT = 0>
__global__ void __launch_bounds__(MAXT, MINB)
align_fused_kernel(const mucon_viterbi_batch b, const int J, const BST* __restrict__ logp,
                   const int32_t* __restrict__ order, const FusedCfg cfg) {
  extern __shared__ __align__(128) unsigned char sm[];
  // Let a programmatically dependent grid (the main launch that follows the long-tail launch in the
  // same stream, mucon_viterbi_align_fused_tail) start now; a no-op otherwise.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int fs = FS ? FS : b.fs;
  const int C = CT ? CT : b.C;  // CT: class count known at compile time (immediate offsets in the scan)
  const int u = order ? order[blockIdx.x] : blockIdx.x;
  const int v = b.unit_vid[u];
  const int64_t r0 = b.vid_off[v];
  const int64_t T = b.vid_off[v + 1] - r0;
  const int K = static_cast<int>(T / fs);
  const int tr0 = b.tr_off[u];
  const int N = b.tr_off[u + 1] - tr0;
  const bool feasible = K >= 1 && N >= 1 && static_cast<int64_t>(K) <= static_cast<int64_t>(N) * J;

  uint64_t* tma_full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* ring_full = tma_full + 8;
  uint64_t* ring_empty = tma_full + 16;
  BST* slabs = reinterpret_cast<BST*>(sm + kFusedBarBytes);
  const uint32_t slab_elems = static_cast<uint32_t>(cfg.bps) * fs * C;
  BST* ring = slabs + static_cast<size_t>(cfg.stages) * slab_elems;
  size_t dp_off = kFusedBarBytes + (static_cast<size_t>(cfg.stages) * slab_elems +
                                    static_cast<size_t>(cfg.ring_slabs) * cfg.bps * C) * sizeof(BST);
  dp_off = (dp_off + 15) & ~size_t(15);

  if (threadIdx.x == 0) {
    for (int s = 0; s < cfg.stages; ++s) mbar_init(&tma_full[s], 1);
    for (int s = 0; s < cfg.ring_slabs; ++s) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();

  if (static_cast<int>(threadIdx.x) < cfg.scan_threads) {
    // ===================================== scan role =====================================
    if (K < 1) return;
    const int bps = cfg.bps, stages = cfg.stages;
    const uint32_t row_bytes = static_cast<uint32_t>(C) * sizeof(BST);
    const int nslabs = (K + bps - 1) / bps;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(logp) + r0 * row_bytes;
    auto issue = [&](int slab, int st) {
      const int b0 = slab * bps;
      const int nb = min(bps, K - b0);
      const uint32_t bytes = static_cast<uint32_t>(nb) * fs * row_bytes;
      mbar_arrive_expect_tx(&tma_full[st], bytes);
      bulk_g2s(slabs + static_cast<size_t>(st) * slab_elems, src + static_cast<int64_t>(b0) * fs * row_bytes, bytes,
               &tma_full[st]);
    };
    if (threadIdx.x == 0 && cfg.z_off == nullptr) {
      const int pre = min(stages, nslabs);
      for (int s = 0; s < pre; ++s) issue(s, s);
    }
    const int c = threadIdx.x * CPT;
    const bool active = c < C;
    BST run[CPT], prev[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) { run[j] = neg_zero<BST>(); prev[j] = 0; }  // -0: exact additive identity
    BST* out_g = (cfg.write_bs && b.bs) ? const_cast<BST*>(reinterpret_cast<const BST*>(b.bs)) + b.blk_off[v] * C + c
                                        : nullptr;
    int st = 0, rs = 0;
    uint32_t parity = 0, rpass = 0;
    if (cfg.z_off != nullptr) {
      // ---- pooled source: every frame's row is looked up in the [Tz, C] table (L1 / L2 resident: a video's table
      // is at most a few hundred KB and each row serves T/Tz consecutive frames)
      const int64_t z0 = cfg.z_off[v];
      const int Tz = static_cast<int>(cfg.z_off[v + 1] - z0);
      const float scale = static_cast<float>(Tz) / static_cast<float>(T);
      const BST* tab = logp + z0 * C + (active ? c : 0);
      constexpr int CH = 10;  // frames whose loads are in flight together
      for (int i = 0; i < nslabs; ++i) {
        if (feasible && rpass > 0) mbar_wait(&ring_empty[rs], (rpass & 1) ^ 1);
        if (active) {
          const int b0 = i * bps;
          const int nb = min(bps, K - b0);
          BST* rp = ring + static_cast<size_t>(rs) * bps * C + c;
          for (int bb = 0; bb < nb; ++bb) {
            const int t0 = (b0 + bb) * fs;
            for (int r0 = 0; r0 < fs; r0 += CH) {
              BST x[CH][CPT];
#pragma unroll
              for (int r = 0; r < CH; ++r) {
                if (r0 + r < fs) {
                  int iz = static_cast<int>(floorf(static_cast<float>(t0 + r0 + r) * scale));
                  iz = iz > Tz - 1 ? Tz - 1 : iz;
                  const BST* row = tab + static_cast<int64_t>(iz) * C;
                  if constexpr (CPT == 2 && sizeof(BST) == 4) {
                    const float2 xv = __ldg(reinterpret_cast<const float2*>(row));
                    x[r][0] = xv.x;
                    x[r][1] = xv.y;
                  } else {
#pragma unroll
                    for (int j = 0; j < CPT; ++j) x[r][j] = __ldg(row + j);
                  }
                }
              }
#pragma unroll
              for (int r = 0; r < CH; ++r)
                if (r0 + r < fs) {
#pragma unroll
                  for (int j = 0; j < CPT; ++j) run[j] = run[j] + x[r][j];
                }
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
              const BST o = (b0 + bb == 0) ? run[j] : run[j] - prev[j];
              prev[j] = run[j];
              rp[bb * C + j] = o;
              if (out_g) out_g[static_cast<int64_t>(b0 + bb) * C + j] = o;
            }
          }
        }
        if (cfg.scan_threads == 32) __syncwarp(); else named_bar_sync(15, cfg.scan_threads);
        if (threadIdx.x == 0 && feasible) mbar_arrive(&ring_full[rs]);
        if (++rs == cfg.ring_slabs) { rs = 0; ++rpass; }
      }
      return;
    }
    for (int i = 0; i < nslabs; ++i) {
      mbar_wait(&tma_full[st], parity);
      if (feasible && rpass > 0) mbar_wait(&ring_empty[rs], (rpass & 1) ^ 1);  // DP is done with this ring slab
      if (active) {
        const int b0 = i * bps;
        const int nb = min(bps, K - b0);
        const BST* s = slabs + static_cast<size_t>(st) * slab_elems + c;
        BST* rp = ring + static_cast<size_t>(rs) * bps * C + c;
        for (int bb = 0; bb < nb; ++bb) {
          if constexpr (FS != 0 && CPT == 2 && sizeof(BST) == 4) {
            // two adjacent class columns per lane: one 8-byte shared load and one packed
            // add.rn.f32x2 (two independent IEEE float32 additions) per frame
            unsigned long long acc;
            asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(run[0]), "f"(run[1]));
#pragma unroll
            for (int r = 0; r < (FS ? FS : 1); ++r) {
              const unsigned long long x = *reinterpret_cast<const unsigned long long*>(s + r * C);
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(x));
            }
            asm("mov.b64 {%0, %1}, %2;" : "=f"(run[0]), "=f"(run[1]) : "l"(acc));
          } else if (FS) {
#pragma unroll
            for (int r = 0; r < (FS ? FS : 1); ++r)
#pragma unroll
              for (int j = 0; j < CPT; ++j) run[j] = run[j] + s[r * C + j];
          } else {
#pragma unroll 4
            for (int r = 0; r < fs; ++r)
#pragma unroll
              for (int j = 0; j < CPT; ++j) run[j] = run[j] + s[r * C + j];
          }
          s += fs * C;
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const BST o = (b0 + bb == 0) ? run[j] : run[j] - prev[j];
            prev[j] = run[j];
            rp[bb * C + j] = o;
            if (out_g) out_g[static_cast<int64_t>(b0 + bb) * C + j] = o;
          }
        }
      }
      if (cfg.scan_threads == 32) __syncwarp(); else named_bar_sync(15, cfg.scan_threads);  // stage consumed, ring slab written
      if (threadIdx.x == 0) {
        if (feasible) mbar_arrive(&ring_full[rs]);
        if (i + stages < nslabs) issue(i + stages, st);
      }
      if (++st == stages) { st = 0; parity ^= 1; }
      if (++rs == cfg.ring_slabs) { rs = 0; ++rpass; }
    }
    return;
  }

  // ======================================= DP role =======================================
  constexpr int kSegsPerWarp = 32 / G;
  const int dpt = threadIdx.x - cfg.scan_threads;
  DpTeam t;
  t.u = u;
  t.lane = dpt & 31;
  t.wl = dpt >> 5;
  t.nw = max(1, (N - 1 + kSegsPerWarp - 1) / kSegsPerWarp);
  if (t.wl >= t.nw) return;  // this unit needs fewer warps than the CTA has
  t.ltid = dpt;
  t.nthr = t.nw * 32;
  t.bar_id = 1;
  t.slot = 0;
  t.c0 = 0;
  t.NS = dp_columns(cfg.dp_warps, G);
  if (!dp_feasible(b, J, t)) return;

  const DpLayout L = dp_layout(cfg.dp_warps, G, J, 0, cfg.bp_rows);
  unsigned char* dsm = sm + dp_off;
  DpShared sh;
  sh.rows0 = reinterpret_cast<double*>(dsm + L.rows0);
  sh.Ex = reinterpret_cast<double*>(dsm + L.Ex);
  sh.ex_stride = cfg.dp_warps;
  sh.segend = reinterpret_cast<int64_t*>(dsm + L.segend);
  sh.trl = reinterpret_cast<int*>(dsm + L.trl);
  sh.segb = reinterpret_cast<int*>(dsm + L.segb);
  sh.fin_v = reinterpret_cast<double*>(dsm + L.fin_v);
  sh.fin_j = reinterpret_cast<int*>(dsm + L.fin_j);
  sh.bpS = dsm + L.bpS;
  sh.bp_rows = cfg.bp_rows;
  for (int n = t.ltid; n < N; n += t.nthr) sh.trl[n] = b.tr[tr0 + n];
  t.sync();  // trl visible

  RingSrc<BST> src;
  src.ring = ring;
  src.full = ring_full;
  src.empty = ring_empty;
  src.C = C;
  src.bps = cfg.bps;
  src.slabs = cfg.ring_slabs;
  dp_unit<BST, G, SL>(b, J, t, sh, src);
}

}  // namespace mucon
