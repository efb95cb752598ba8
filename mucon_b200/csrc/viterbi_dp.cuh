// viterbi_dp.cuh -- the K x N x J dynamic program, traceback and label writer (sm_100a).
//
// Follows reference src/core/viterbi/viterbi.py:81-158 in the dense form of SURVEY.md section 8a:
//   step k, segment n, age j (length in blocks before this step):
//     a = S[n][j] + bs_k[tr_n]                       stay, lands at age j+1   (viterbi.py:97-104)
//     c = (a + rows[n][j]) + 0.0                     advance candidate        (viterbi.py:106-121)
//     S'[n+1][1] = fold_j c, ascending j, replace iff old <= new             (viterbi.py:26-28)
//     bp[k][n+1] = winning j
//
// Work decomposition
//   unit     one (video, candidate transcript).
//   segment 0 has exactly one hypothesis (entered at step 0), so it is a scalar running sum kept
//            by lane 0 of the unit's first warp -- in float32 when the reference's NumPy would keep
//            it in float32, float64 otherwise.
//   segment n >= 1 is a shift register over ages 1..J held by a group of 8 lanes, SL = ceil(J/8)
//            consecutive ages per lane, all in registers with static indices: ageing is the
//            in-place update R[i] = R[i-1] + b, the value leaving a lane moves to the next lane
//            with one shuffle, and the same shuffle hands the group's winner to the next segment's
//            group.  Four segments per warp; a unit with N <= 5 is a single warp with no block
//            barrier at all, larger units use ceil((N-1)/4) warps and one named barrier per step.
//   dead     hypotheses are -inf: no liveness flags.  The only place liveness is observable is a
//            fold whose maximum is -inf; there the winner is the oldest live age, which follows
//            from (k, n, J) alone:  live ages at step k are [max(1, k-n*J), min(J, k-n)].
//   CTA      a bin of units packed by the host (mucon_viterbi_pack_h); every unit synchronises on
//            its own named barrier, units in a bin never wait for each other.
//   bp       staged in shared memory for the traceback, flushed to HBM once per unit.
#pragma once
#include <math.h>

#include "common.cuh"

namespace mucon {

constexpr int kDpMaxWarps = 16;  // warps per CTA: 4, 8 or 16 (chosen by mucon_viterbi_pack_h)
constexpr int kDpChunk = 32;     // DP steps per block-score staging chunk
constexpr int kDpMaxJ = 128;     // ages live in registers: 4 lanes x <= 17 or 8 lanes x <= 16

// lanes per segment for a given J: 4 (8 segments per warp) while the per-lane register file
// holds the ages, else 8
__host__ __device__ inline int dp_group(int J) { return (J + 3) / 4 <= 17 ? 4 : 8; }
__host__ __device__ inline int dp_max_n(int G) { return 1 + kDpMaxWarps * (32 / G); }

// Order-preserving map double -> uint64 (and back).  FP64 compares sit on a long-latency pipe and
// the fold is the loop-carried critical path, so all arg-max work is done on integer keys.  Equal
// doubles have equal keys except -0.0 / +0.0; candidates are a + row with row != -0.0 (rows are
// canonicalised when loaded), and such a sum is never -0.0, which also makes the reference's
// trailing "+ 0.0" (viterbi.py:116) the identity.
__device__ __forceinline__ unsigned long long dkey(double c) {
  const long long b = __double_as_longlong(c);
  return static_cast<unsigned long long>(b) ^ (static_cast<unsigned long long>(b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
  return __longlong_as_double(static_cast<long long>(b));
}
constexpr unsigned long long kKeyNegInf = 0x000fffffffffffffull;  // dkey(-inf)

__device__ __forceinline__ int label_of_frame(int64_t t, int64_t rem, const int32_t* trl, const int64_t* segend,
                                              int last) {
  if (t < rem) return trl[last];
  int n = 0;
  while (n < last && t >= segend[n]) ++n;
  return trl[n];
}

// Writes T labels at out with `nth` cooperating threads (this thread is `tid`).
// segend[n] = rem + fs * sum_{m<=n} blocks[m] (exclusive end, frames).
__device__ __forceinline__ void write_labels(int32_t* out, int64_t T, int64_t rem, const int32_t* trl,
                                             const int64_t* segend, int last, int tid, int nth) {
  const int64_t mis = (reinterpret_cast<uintptr_t>(out) >> 2) & 3;
  int64_t head = (4 - mis) & 3;
  if (head > T) head = T;
  for (int64_t t = tid; t < head; t += nth) out[t] = label_of_frame(t, rem, trl, segend, last);
  const int64_t nvec = (T - head) >> 2;
  int4* out4 = reinterpret_cast<int4*>(out + head);
  for (int64_t q = tid; q < nvec; q += nth) {
    const int64_t t = head + 4 * q;
    int lab[4];
    if (t + 3 < rem) {
      lab[0] = lab[1] = lab[2] = lab[3] = trl[last];
    } else {
      int n = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t te = t + e;
        if (te < rem) {
          lab[e] = trl[last];
        } else {
          while (n < last && te >= segend[n]) ++n;
          lab[e] = trl[n];
        }
      }
    }
    out4[q] = make_int4(lab[0], lab[1], lab[2], lab[3]);
  }
  for (int64_t t = head + 4 * nvec + tid; t < T; t += nth) out[t] = label_of_frame(t, rem, trl, segend, last);
}

// Shared-memory plan of one CTA of `wpc` warps.  Every warp owns 32/G + 1 segment columns (its
// segments plus a spare for segment 0 of a unit starting there), so units never share a column:
// segment n of the unit whose first warp is w0 lives in column (32/G + 1)*w0 + n.
__host__ __device__ inline int dp_columns(int wpc, int G) { return wpc * (32 / G + 1); }

struct DpLayout {
  size_t rows0, Ex, segend, trl, segb, fin_v, fin_j, bsS, bpS, total;
};
__host__ __device__ inline DpLayout dp_layout(int wpc, int G, int J, int bs_elem, int bp_rows) {
  const int kDpNS = dp_columns(wpc, G), kDpWarps = wpc;
  DpLayout L;
  size_t o = 0;
  L.rows0 = o; o += sizeof(double) * (size_t)kDpWarps * J;  // segment-0 length rows, one per unit slot
  L.Ex = o; o += sizeof(double) * 2 * kDpWarps;             // cross-warp entry scores by step parity
  L.segend = o; o += sizeof(int64_t) * kDpNS;
  L.fin_v = o; o += sizeof(double) * kDpWarps;
  L.trl = o; o += sizeof(int) * kDpNS;
  L.segb = o; o += sizeof(int) * kDpNS;
  L.fin_j = o; o += sizeof(int) * kDpWarps;
  o = (o + 15) & ~size_t(15);
  L.bsS = o; o += (size_t)bs_elem * 2 * kDpChunk * kDpNS;
  o = (o + 15) & ~size_t(15);
  L.bpS = o; o += (size_t)bp_rows * kDpNS;
  L.total = (o + 15) & ~size_t(15);
  return L;
}

// length score of `age` blocks for the label with parameters g[0..2]: ((l*ln m - m) - lf_l) - norms
// (length_model.py:65-71), -inf when the length is not representable (length_model.py:76-80).
__device__ __forceinline__ double length_row(const mucon_viterbi_batch& b, int tr0, int n, int age, int J) {
  if (age < 1 || age > J) return -INFINITY;
  if (b.len_rows) return __dadd_rn(b.len_rows[static_cast<size_t>(tr0 + n) * J + age - 1], 0.0);
  const int l = age * b.fs;
  if (l >= b.max_len) return -INFINITY;
  const double* g = b.len_params + static_cast<size_t>(tr0 + n) * 3;
  double r = __dmul_rn(static_cast<double>(l), g[0]);
  r = __dsub_rn(r, g[1]);
  r = __dsub_rn(r, b.logfact[age]);
  r = __dsub_rn(r, g[2]);
  return __dadd_rn(r, 0.0);
}

template <typename BST, int G, int SL>
__global__ void __launch_bounds__(kDpMaxWarps * 32, 1)
dp_kernel(const mucon_viterbi_batch b, const int J, const int32_t* __restrict__ warp_unit, const int bp_rows) {
  extern __shared__ __align__(16) unsigned char sm[];
  constexpr int kSegsPerWarp = 32 / G;
  const int kDpWarps = blockDim.x >> 5, kDpNS = dp_columns(kDpWarps, G);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* wu = warp_unit + static_cast<size_t>(blockIdx.x) * kDpWarps;
  const int u = wu[warp];
  if (u < 0) return;  // unused warp of a partially filled bin
  int w0 = warp, w1 = warp + 1;
  while (w0 > 0 && wu[w0 - 1] == u) --w0;
  while (w1 < kDpWarps && wu[w1] == u) ++w1;
  const int nw = w1 - w0, wl = warp - w0;
  const int ltid = wl * 32 + lane, nthr = nw * 32;
  const int c0 = w0 * (kSegsPerWarp + 1);  // first shared-memory column of this unit
  auto ubar = [&]() {
    if (nw == 1) __syncwarp(); else named_bar_sync(1 + w0, nthr);
  };

  const int v = b.unit_vid[u];
  const int64_t T = b.vid_off[v + 1] - b.vid_off[v];
  const int fs = b.fs;
  const int K = static_cast<int>(T / fs);
  const int tr0 = b.tr_off[u];
  const int N = b.tr_off[u + 1] - tr0;
  const int C = b.C;

  if (K < 1 || N < 1 || static_cast<int64_t>(K) > static_cast<int64_t>(N) * J) {
    if (ltid == 0) {
      b.status[u] = MUCON_UNIT_INFEASIBLE;
      b.score[u] = __longlong_as_double(0x7ff8000000000000ll);
      b.final_j[u] = 0;
    }
    for (int n = ltid; n < N; n += nthr) b.seg_blocks[tr0 + n] = 0;
    return;
  }

  const DpLayout L = dp_layout(kDpWarps, G, J, sizeof(BST), bp_rows);
  double* rows0 = reinterpret_cast<double*>(sm + L.rows0) + static_cast<size_t>(w0) * J;  // [J], ages 1..J
  double* Ex = reinterpret_cast<double*>(sm + L.Ex);                                       // [2][16] by warp
  int64_t* segend = reinterpret_cast<int64_t*>(sm + L.segend) + c0;
  int* trl = reinterpret_cast<int*>(sm + L.trl) + c0;
  int* segb = reinterpret_cast<int*>(sm + L.segb) + c0;
  double* fin_v = reinterpret_cast<double*>(sm + L.fin_v) + w0;
  int* fin_j = reinterpret_cast<int*>(sm + L.fin_j) + w0;
  BST* bsS = reinterpret_cast<BST*>(sm + L.bsS);
  uint8_t* bpS = sm + L.bpS;
  const BST* bs_g = reinterpret_cast<const BST*>(b.bs) + b.blk_off[v] * C;
  uint8_t* bp_g = b.bp + b.bp_off[u];
  const bool bp_in_smem = bp_rows >= K;

  for (int n = ltid; n < N; n += nthr) trl[n] = b.tr[tr0 + n];
  for (int j = ltid; j < J; j += nthr) rows0[j] = length_row(b, tr0, 0, j + 1, J);

  const int64_t rem = T - static_cast<int64_t>(K) * fs;
  int last = N - 1;

  if (K < N) {
    // Nothing reaches the last segment: the reference returns -inf and the path with one block
    // in each of the first K segments (viterbi.py:125-138).
    last = K - 1;
    for (int n = ltid; n < N; n += nthr) segb[n] = (n < K) ? 1 : 0;
    if (ltid == 0) {
      b.status[u] = MUCON_UNIT_SHORT;
      b.score[u] = -INFINITY;
      b.final_j[u] = 1;
    }
    for (int i = ltid; i < K * N; i += nthr) bp_g[i] = 0;  // not computed
    ubar();
  } else {
    ubar();  // trl visible
    auto stage = [&](int chunk) {
      const int k0 = chunk * kDpChunk;
      const int nk = min(kDpChunk, K - k0);
      BST* dst = bsS + static_cast<size_t>(chunk & 1) * kDpChunk * kDpNS + c0;
      for (int i = ltid; i < nk * N; i += nthr) {
        const int kk = i / N, n = i - kk * N;
        const BST* src = bs_g + static_cast<int64_t>(k0 + kk) * C + trl[n];
        if (sizeof(BST) == 4) cp_async4(dst + kk * kDpNS + n, src); else cp_async8(dst + kk * kDpNS + n, src);
      }
      cp_async_commit();
    };
    const int nchunks = (K + kDpChunk - 1) / kDpChunk;
    stage(0);
    if (nchunks > 1) stage(1);

    // this lane's segment (n >= 1), its ages a0+1 .. a0+SL and their length scores
    const int g = lane / G, lig = lane - g * G;
    const int n = 1 + wl * kSegsPerWarp + g;
    const bool has_seg = n < N;
    const int a0 = lig * SL;
    double R[SL], rowr[SL];
#pragma unroll
    for (int i = 0; i < SL; ++i) {
      R[i] = -INFINITY;
      rowr[i] = has_seg ? length_row(b, tr0, n, a0 + i + 1, J) : -INFINITY;
    }
    const int nJ = (n <= 0x7fffffff / J) ? n * J : 0x7fffffff;

    if (nchunks > 1) cp_async_wait<1>(); else cp_async_wait<0>();
    ubar();

    // segment 0: scalar chain on lane 0 of the first warp; 0.0 + F[fs-1, tr_0]  (viterbi.py:81-90)
    const bool f32seg0 = (sizeof(BST) == 4) && b.seg0_f32;
    double s0 = __dadd_rn(0.0, static_cast<double>(bsS[c0]));

    int kk = 1, chunk = 0;
    for (int k = 1; k < K; ++k) {
      if (kk == kDpChunk) {
        kk = 0;
        ++chunk;
        cp_async_wait<0>();
        ubar();  // chunk landed for every thread of the unit; the other buffer is free
        if (chunk + 1 < nchunks) stage(chunk + 1);
      }
      const BST* bsk = bsS + (static_cast<size_t>(chunk & 1) * kDpChunk + kk) * kDpNS + c0;
      const int par = k & 1;

      // ---- segment 0 -> entry of segment 1 (uniform work, only lane 0 of warp 0 keeps the result)
      double e1 = -INFINITY;
      if (wl == 0) {
        const BST b0 = bsk[0];
        const double a = f32seg0 ? static_cast<double>(__fadd_rn(static_cast<float>(s0), static_cast<float>(b0)))
                                 : __dadd_rn(s0, static_cast<double>(b0));
        s0 = a;
        // the single hypothesis of segment 0 has age k; it can advance while k <= J
        e1 = (k <= J) ? __dadd_rn(a, rows0[k - 1]) : -INFINITY;  // rows are never -0.0: no "+ 0.0" needed
      }

      // ---- segments >= 1: age every hypothesis, build the advance candidates, fold
      const double bd = static_cast<double>(bsk[has_seg ? n : 0]);
      const double out = __dadd_rn(R[SL - 1], bd);  // leaves this lane
#pragma unroll
      for (int i = SL - 1; i >= 1; --i) R[i] = __dadd_rn(R[i - 1], bd);
      // Candidates: old position i (age a0+i+1) now sits at R[i+1] / out.  The fold is a max by
      // (value, age) on integer keys.  Positions 1.. do not depend on this step's entry: they are
      // reduced by a tree (older position wins ties); the youngest candidate -- the only one on
      // the loop-carried path -- joins last and wins only when strictly greater.
      unsigned long long key[SL];
      int idx[SL];
#pragma unroll
      for (int i = 0; i < SL; ++i) {
        const double a = (i + 1 < SL) ? R[(i + 1 < SL) ? i + 1 : 0] : out;
        key[i] = dkey(__dadd_rn(a, rowr[i]));
        idx[i] = i;
      }
#pragma unroll
      for (int w = 1; w < SL - 1; w <<= 1) {
#pragma unroll
        for (int i = 1; i + w < SL; i += 2 * w) {
          if (key[i + w] >= key[i]) { key[i] = key[i + w]; idx[i] = idx[i + w]; }
        }
      }
      unsigned long long bk = key[0];
      int bi = 0;
      if (SL > 1 && key[1] >= bk) { bk = key[1]; bi = idx[1]; }
      int bage = a0 + bi + 1;
      // butterfly over the group's lanes: the partner with the higher lane holds older ages, so
      // it wins ties; the partner with the lower lane must be strictly greater
#pragma unroll
      for (int off = 1; off < G; off <<= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, off);
        const int oa = __shfl_xor_sync(0xffffffffu, bage, off);
        const bool take = (lane & off) ? (ok > bk) : (ok >= bk);
        if (take) { bk = ok; bage = oa; }
      }
      const double bv = dunkey(bk);
      // a fold whose maximum is -inf is decided by liveness alone: oldest live age, or no entry
      {
        const int jhi = min(J, k - n), jlo = max(1, k - nJ);
        if (bk == kKeyNegInf) bage = (jlo <= jhi) ? jhi : 0;
      }
      // shift between lanes; the last lane of a group forwards the group's winner instead
      const double send = (lig == G - 1) ? bv : out;
      double inc = __shfl_up_sync(0xffffffffu, send, 1);
      if (lane == 0) inc = (wl == 0) ? e1 : -INFINITY;  // warp 0: from segment 0; others: patched below
      if (has_seg && lig == G - 1 && n + 1 < N) {
        if (bp_in_smem) bpS[static_cast<size_t>(k) * kDpNS + c0 + n + 1] = static_cast<uint8_t>(bage);
        else bp_g[static_cast<int64_t>(k) * N + n + 1] = static_cast<uint8_t>(bage);
      }
      if (ltid == 0 && N > 1) {
        const uint8_t j1 = (k <= J) ? static_cast<uint8_t>(k) : uint8_t(0);
        if (bp_in_smem) bpS[static_cast<size_t>(k) * kDpNS + c0 + 1] = j1;
        else bp_g[static_cast<int64_t>(k) * N + 1] = j1;
      }
      if (nw > 1) {
        if (lane == 31 && wl + 1 < nw) Ex[par * kDpWarps + warp + 1] = bv;
        named_bar_sync(1 + w0, nthr);
        if (lane == 0 && wl > 0) inc = Ex[par * kDpWarps + warp];
      }
      R[0] = inc;
      ++kk;
    }

    // end symbol: fold over the last segment (viterbi.py:125-138)
    if (N == 1) {
      if (ltid == 0) {
        *fin_v = __dadd_rn(__dadd_rn(s0, rows0[K - 1]), 0.0);  // K <= J is guaranteed by feasibility
        *fin_j = K;
      }
    } else if (n == N - 1) {
      double bv = -INFINITY;
      int bi = 0;
#pragma unroll
      for (int i = 0; i < SL; ++i) {
        const double c = __dadd_rn(R[i], rowr[i]);
        if (i == 0 || c >= bv) { bv = c; bi = i; }
      }
      int bage = a0 + bi + 1;
#pragma unroll
      for (int off = 1; off < G; off <<= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu >> (32 - G) << (g * G), bv, off);
        const int oa = __shfl_xor_sync(0xffffffffu >> (32 - G) << (g * G), bage, off);
        const bool take = (lane & off) ? (ov > bv) : (ov >= bv);
        if (take) { bv = ov; bage = oa; }
      }
      bv = __dadd_rn(bv, 0.0);
      // after the last step (k = K-1) the live ages of segment n are [max(1, K-n*J), min(J, K-n)]
      const int jhi = min(J, K - n), jlo = max(1, K - nJ);
      if (bv == -INFINITY) bage = (jlo <= jhi) ? jhi : 0;
      if (lig == 0) { *fin_v = bv; *fin_j = bage; }
    }
    ubar();
    if (ltid == 0) {  // traceback over the back-pointer table (viterbi.py:140-153)
      const double sc = *fin_v;
      const int jf = *fin_j;
      int m = N - 1;
      int k0 = K - jf;
      segb[m] = jf;
      while (m > 0) {
        const int ln = bp_in_smem ? static_cast<int>(bpS[static_cast<size_t>(k0) * kDpNS + c0 + m])
                                  : static_cast<int>(__ldcg(bp_g + static_cast<int64_t>(k0) * N + m));
        segb[m - 1] = ln;
        k0 -= ln;
        --m;
      }
      b.score[u] = sc;
      b.final_j[u] = jf;
      b.status[u] = (isfinite(sc) || sc == -INFINITY) ? MUCON_UNIT_OK : MUCON_UNIT_NONFINITE;
    }
    ubar();
    // back-pointer table -> HBM, [K, N] row-major; row 0 and column 0 hold no entries
    if (bp_in_smem) {
      const int total = K * N;
      int k = ltid / N, m = ltid - k * N;
      const int dk = nthr / N, dn = nthr - dk * N;
      for (int i = ltid; i < total; i += nthr) {
        bp_g[i] = (k > 0 && m > 0) ? bpS[static_cast<size_t>(k) * kDpNS + c0 + m] : uint8_t(0);
        k += dk;
        m += dn;
        if (m >= N) { m -= N; ++k; }
      }
    } else {
      for (int i = ltid; i < N; i += nthr) bp_g[i] = 0;
      for (int k = 1 + ltid; k < K; k += nthr) bp_g[static_cast<int64_t>(k) * N] = 0;
    }
  }

  for (int m = ltid; m < N; m += nthr) b.seg_blocks[tr0 + m] = segb[m];
  const int64_t lo = b.lab_off ? b.lab_off[u] : -1;
  if (lo >= 0) {
    if (ltid == 0) {
      int64_t pos = rem;
      for (int m = 0; m < N; ++m) { pos += static_cast<int64_t>(fs) * segb[m]; segend[m] = pos; }
    }
    ubar();
    write_labels(b.labels + lo, T, rem, trl, segend, last, ltid, nthr);
  }
}

}  // namespace mucon
