// viterbi_lanes.cuh -- the K x N x J dynamic program with one LANE per transcript segment.
//
// Same program as dp_unit (viterbi_dp.cuh; reference src/core/viterbi/viterbi.py:81-158), different
// decomposition.  dp_unit spreads the J ages of a segment over 8 lanes, so every step pays a
// three-level shuffle butterfly, asymmetric tie logic and a lane shift: ~56 warp instructions per
// segment-step, most of them on the loop-carried critical path.  Here a lane owns a whole segment:
//   * its J hypothesis scores are a shift register in the lane's own registers (static indices),
//   * the fold over the J candidates is an in-lane tree -- no shuffles, massive ILP,
//   * the only cross-lane traffic per step is ONE shuffle handing the winner to the next segment,
// ~15 warp instructions per segment-step.  A warp carries 32 segments, i.e. several units
// (video x transcript) of similar length packed side by side by mucon_viterbi_pack_lanes_h; units
// never straddle warps, warps never synchronise with each other.
//
// Block scores are read from HBM/L2 (written by the scan kernel, which precedes this kernel in the stream).
#pragma once
#include <math.h>

#include "viterbi_dp.cuh"

namespace mucon {

constexpr int kLanesChunk = 8;   // DP steps per block-score staging chunk

struct LanesLayout {
  size_t rows, stage, trl, segend, total;
};
__host__ __device__ inline LanesLayout lanes_layout(int JT, int bs_elem) {
  LanesLayout L;
  size_t o = 0;
  L.rows = o; o += sizeof(double) * (size_t)JT * 32;
  L.stage = o; o += (size_t)bs_elem * 2 * kLanesChunk * 2 * 32;
  o = (o + 15) & ~size_t(15);
  L.segend = o; o += sizeof(int64_t) * 40;
  L.trl = o; o += sizeof(int) * 40;
  L.total = (o + 15) & ~size_t(15);
  return L;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// CH ages per fold chunk, NCH chunks: JT = CH * NCH >= J register positions per lane.
template <typename BST, int CH, int NCH>
__global__ void __launch_bounds__(32, 1)
dp_lanes_kernel(const mucon_viterbi_batch b, const int J, const int32_t* __restrict__ lane_unit,
                const int* __restrict__ /*reserved*/) {
  constexpr int JT = CH * NCH;
  static_assert(CH % 2 == 0, "rows are read two at a time");
  extern __shared__ __align__(16) unsigned char sm[];
  const LanesLayout L = lanes_layout(JT, sizeof(BST));
  double2* rows2 = reinterpret_cast<double2*>(sm + L.rows);  // [JT/2][32]
  BST* stg = reinterpret_cast<BST*>(sm + L.stage);           // [2][chunk][2][32]
  int64_t* segend = reinterpret_cast<int64_t*>(sm + L.segend);
  int* trl_s = reinterpret_cast<int*>(sm + L.trl);

  const int lane = threadIdx.x;
  const unsigned full = 0xffffffffu;
  const int u = lane_unit[static_cast<size_t>(blockIdx.x) * 32 + lane];
  const bool active = u >= 0;
  const unsigned same = __match_any_sync(full, u);
  const int first = __ffs(same) - 1;
  const bool is_first = active && lane == first;
  const int n = 1 + lane - first;  // this lane's segment (segment 0 rides on the first lane)

  int v = 0, tr0 = 0, N = 0, fs = b.fs;
  int64_t T = 0;
  int K = 0;
  if (active) {
    v = b.unit_vid[u];
    T = b.vid_off[v + 1] - b.vid_off[v];
    K = static_cast<int>(T / fs);
    tr0 = b.tr_off[u];
    N = b.tr_off[u + 1] - tr0;
  }
  const bool feasible = active && K >= 1 && N >= 1 && static_cast<int64_t>(K) <= static_cast<int64_t>(N) * J;
  const bool is_short = feasible && K < N;
  const bool run = feasible && !is_short;
  const bool has_seg = run && n < N;
  const int C = b.C;
  uint8_t* bp_g = active ? b.bp + b.bp_off[u] : nullptr;

  if (active && !feasible && is_first) {
    b.status[u] = MUCON_UNIT_INFEASIBLE;
    put_score(b, u, __longlong_as_double(0x7ff8000000000000ll));
    b.final_j[u] = 0;
    for (int m = 0; m < N; ++m) put_seg(b, tr0 + m, 0);
  }
  if (is_short && is_first) {
    // nothing reaches the last segment (viterbi.py:125-138)
    b.status[u] = MUCON_UNIT_SHORT;
    put_score(b, u, -INFINITY);
    b.final_j[u] = 1;
    for (int m = 0; m < N; ++m) put_seg(b, tr0 + m, (m < K) ? 1 : 0);
    for (int i = 0; i < K * N; ++i) bp_g[i] = 0;
  }
  if (run && is_first) {  // row 0 and column 0 of the back-pointer table hold no entries
    for (int m = 0; m < N; ++m) bp_g[m] = 0;
    for (int k = 1; k < K; ++k) bp_g[static_cast<int64_t>(k) * N] = 0;
  }

  // length scores of this lane's segment, ages 1..JT (beyond J: -inf)
  {
    double* rows = reinterpret_cast<double*>(rows2);
#pragma unroll 1
    for (int i = 0; i < JT; ++i)
      rows[(static_cast<size_t>(i >> 1) * 32 + lane) * 2 + (i & 1)] =
          has_seg ? length_row(b, tr0, n, i + 1, J) : -INFINITY;
  }
  const int col = run ? b.tr[tr0 + (has_seg ? n : 0)] : 0;
  const int col0 = run ? b.tr[tr0] : 0;
  const BST* bs_v = reinterpret_cast<const BST*>(b.bs) + (run ? b.blk_off[v] * C : 0);
  const int Kmax = __reduce_max_sync(full, run ? K : 0);
  const int nJ = (n <= 0x7fffffff / J) ? n * J : 0x7fffffff;
  const bool f32seg0 = (sizeof(BST) == 4) && b.seg0_f32;
  const bool bp_writer = has_seg && n + 1 < N;
  const bool bp1_writer = run && is_first && N > 1;

  // ---- block-score staging: chunk ch = steps [ch*kLanesChunk, ...), both columns of this lane
  auto stage = [&](int ch) {
    const int k0 = ch * kLanesChunk;
    if (run && k0 < K) {
      const int k1 = min(K, k0 + kLanesChunk);
      BST* dst = stg + static_cast<size_t>(ch & 1) * kLanesChunk * 64 + lane;
      const BST* src = bs_v + static_cast<int64_t>(k0) * C;
      for (int r = 0; r < k1 - k0; ++r) {
        if (sizeof(BST) == 4) {
          cp_async4(dst + r * 64, src + col);
          cp_async4(dst + r * 64 + 32, src + col0);
        } else {
          cp_async8(dst + r * 64, src + col);
          cp_async8(dst + r * 64 + 32, src + col0);
        }
        src += C;
      }
    }
    cp_async_commit();
  };

  double g0 = 0.0, g1 = 0.0, g2 = 0.0;  // Poisson parameters of segment 0 (ln m, m, norms)
  if (run && b.len_params) {
    g0 = b.len_params[static_cast<size_t>(tr0) * 3];
    g1 = b.len_params[static_cast<size_t>(tr0) * 3 + 1];
    g2 = b.len_params[static_cast<size_t>(tr0) * 3 + 2];
  }
  double R[JT];
#pragma unroll
  for (int i = 0; i < JT; ++i) R[i] = -INFINITY;
  double s0 = 0.0;

  const int nchunks = (Kmax + kLanesChunk - 1) / kLanesChunk;
  if (nchunks > 0) {
    stage(0);
    if (nchunks > 1) { stage(1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
    if (run) s0 = __dadd_rn(0.0, static_cast<double>(stg[32 + lane]));  // 0.0 + F[fs-1, tr_0]
  }

  int chunk = 0, kk = 0;
#pragma unroll 1
  for (int k = 1; k < Kmax; ++k) {
    if (++kk == kLanesChunk) {
      kk = 0;
      ++chunk;
      cp_async_wait<0>();
      __syncwarp();
      if (chunk + 1 < nchunks) stage(chunk + 1);
    }
    const bool live = run && k < K;
    const unsigned lmask = __ballot_sync(full, live);
    if (live) {
      const BST* srow = stg + (static_cast<size_t>(chunk & 1) * kLanesChunk + kk) * 64 + lane;
      const double bd = static_cast<double>(srow[0]);
      const BST b0 = srow[32];
      // segment 0: one hypothesis of age k (viterbi.py:97-121 for n = 0)
      double a0v;
      if (f32seg0) a0v = static_cast<double>(__fadd_rn(static_cast<float>(s0), static_cast<float>(b0)));
      else a0v = __dadd_rn(s0, static_cast<double>(b0));
      s0 = a0v;
      double e1 = -INFINITY;
      if (k <= J) {
        double r0;  // length score of k blocks for segment 0 (= length_row(b, tr0, 0, k, J))
        if (b.len_rows) r0 = __dadd_rn(b.len_rows[static_cast<size_t>(tr0) * J + k - 1], 0.0);
        else if (k * fs >= b.max_len) r0 = -INFINITY;
        else r0 = __dadd_rn(__dsub_rn(__dsub_rn(__dsub_rn(__dmul_rn(static_cast<double>(k * fs), g0), g1),
                                                b.logfact[k]), g2), 0.0);
        e1 = __dadd_rn(a0v, r0);
      }

      // ages, oldest chunk first: a_i = R[i] + b lands at position i+1, candidate a_i + rows[i];
      // an older candidate wins ties ("replace iff old <= new", viterbi.py:26-28)
      double bv = -INFINITY;
      int bi = JT - 1;
#pragma unroll
      for (int c = NCH - 1; c >= 0; --c) {
        double cv[CH];
        int ci[CH];
#pragma unroll
        for (int h = CH / 2 - 1; h >= 0; --h) {
          const double2 rr = rows2[static_cast<size_t>((c * CH) / 2 + h) * 32 + lane];
#pragma unroll
          for (int e = 1; e >= 0; --e) {
            const int i = c * CH + 2 * h + e;
            const double a = __dadd_rn(R[i], bd);
            if (i + 1 < JT) R[(i + 1 < JT) ? i + 1 : 0] = a;
            cv[2 * h + e] = __dadd_rn(a, e ? rr.y : rr.x);
            ci[2 * h + e] = i;
          }
        }
        // in-chunk tree; position p+w is older than p
#pragma unroll
        for (int w = 1; w < CH; w <<= 1) {
#pragma unroll
          for (int p = 0; p + w < CH; p += 2 * w) {
            const bool older = cv[p + w] >= cv[p];
            cv[p] = older ? cv[p + w] : cv[p];
            ci[p] = older ? ci[p + w] : ci[p];
          }
        }
        if (c == NCH - 1) { bv = cv[0]; bi = ci[0]; }
        else {
          const bool younger = cv[0] > bv;
          bv = younger ? cv[0] : bv;
          bi = younger ? ci[0] : bi;
        }
      }
      int bage = bi + 1;
      // a fold whose maximum is -inf is decided by liveness alone (see dp_unit)
      const int jhi = min(J, k - n), jlo = max(1, k - nJ);
      if (bv == -INFINITY) bage = (jlo <= jhi) ? jhi : 0;
      if (bp_writer) bp_g[static_cast<int64_t>(k) * N + n + 1] = static_cast<uint8_t>(bage);
      if (bp1_writer) bp_g[static_cast<int64_t>(k) * N + 1] = (k <= J) ? static_cast<uint8_t>(k) : uint8_t(0);
      double inc = __shfl_up_sync(lmask, bv, 1);
      inc = is_first ? e1 : inc;
      R[0] = inc;
    }
  }
  cp_async_wait<0>();
  __syncwarp();  // back-pointers written by the other lanes are read by the traceback below

  // ---- end symbol: fold over the last segment (viterbi.py:125-138), traceback (:140-153)
  if (run && (n == N - 1 || (N == 1 && is_first))) {
    double bv;
    int bage;
    if (N == 1) {
      bv = __dadd_rn(__dadd_rn(s0, length_row(b, tr0, 0, K, J)), 0.0);  // K <= J by feasibility
      bage = K;
    } else {
      bv = -INFINITY;
      int bi = JT - 1;
#pragma unroll
      for (int i = JT - 1; i >= 0; --i) {
        const double2 rr = rows2[static_cast<size_t>(i >> 1) * 32 + lane];
        const double c = __dadd_rn(R[i], (i & 1) ? rr.y : rr.x);
        if (i == JT - 1 || c > bv) { bv = c; bi = i; }
      }
      bage = bi + 1;
      bv = __dadd_rn(bv, 0.0);
      const int jhi = min(J, K - n), jlo = max(1, K - nJ);
      if (bv == -INFINITY) bage = (jlo <= jhi) ? jhi : 0;
    }
    int m = N - 1;
    int k0 = K - bage;
    put_seg(b, tr0 + m, bage);
    while (m > 0) {
      int ln;
      if (m == 1) ln = (k0 <= J) ? k0 : 0;
      else ln = static_cast<int>(__ldcg(bp_g + static_cast<int64_t>(k0) * N + m));
      put_seg(b, tr0 + m - 1, ln);
      k0 -= ln;
      --m;
    }
    put_score(b, u, bv);
    b.final_j[u] = bage;
    b.status[u] = (isfinite(bv) || bv == -INFINITY) ? MUCON_UNIT_OK : MUCON_UNIT_NONFINITE;
  }
  __syncwarp();

  // ---- labels, one unit at a time with the whole warp
  unsigned todo = __ballot_sync(full, is_first && feasible && b.lab_off && b.lab_off[u] >= 0);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int uu = __shfl_sync(full, u, src);
    const int NN = __shfl_sync(full, N, src);
    const int KK = __shfl_sync(full, K, src);
    const int t0 = __shfl_sync(full, tr0, src);
    const long long TT = __shfl_sync(full, static_cast<long long>(T), src);
    const int64_t rem = TT - static_cast<int64_t>(KK) * fs;
    __syncwarp();
    for (int m = lane; m < NN; m += 32) trl_s[m] = b.tr[t0 + m];
    if (lane == 0) {
      int64_t pos = rem;
      for (int m = 0; m < NN; ++m) {
        pos += static_cast<int64_t>(fs) * __ldcg(b.seg_blocks + t0 + m);
        segend[m] = pos;
      }
    }
    __syncwarp();
    const int last = (KK < NN) ? KK - 1 : NN - 1;
    write_labels(b.labels + b.lab_off[uu], TT, rem, trl_s, segend, last, lane, 32);
  }
}

}  // namespace mucon
