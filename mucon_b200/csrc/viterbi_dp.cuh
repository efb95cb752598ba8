// viterbi_dp.cuh -- the K x N x J dynamic program, traceback and label writer (sm_100a).
//
// Follows reference src/core/viterbi/viterbi.py:81-158 in the dense form of SURVEY.md section 8a.
//
// Work decomposition
//   unit        one (video, candidate transcript); needs ceil(N / SEGS) warps, one warp per SEGS
//               transcript segments.
//   CTA         a bin of units packed by the host (mucon_viterbi_pack_h) so that all 16 warps
//               are busy; every unit synchronises on its own named barrier, units in a bin never
//               wait for each other.
//   warp        lane l owns length slots l, l+32, ... of its segment(s).  A slot is a circular
//               buffer position indexed by (entry step mod J): a hypothesis never moves between
//               lanes, it just ages (len += 1) until len == J, when the slot is recycled for the
//               hypothesis entering at that very step.
//   step k      a   = S + bs_k[tr_n]                            stay        (viterbi.py:97-104)
//               c   = (a + rows[n][len]) + 0.0                  advance     (viterbi.py:106-121)
//               S'[n+1][1] = fold over len ascending, replace iff old <= new (viterbi.py:26-28)
//               The fold is evaluated as: order-preserving 64-bit integer key of c, warp max by
//               two 32-bit REDUX, then the largest len among the slots holding that maximum --
//               the same winner as the sequential fold for every non-NaN input.  Because "+ 0.0"
//               never yields -0.0, equal doubles have equal keys.
//   back-ptrs   bp[k][n] = winning len, staged in shared memory for the traceback and flushed to
//               HBM once per unit with coalesced stores.
#pragma once
#include <math.h>

#include "common.cuh"

namespace mucon {

constexpr int kDpWarps = 16;   // warps per CTA
constexpr int kDpChunk = 32;   // DP steps per block-score staging chunk
constexpr int kDpMaxSlots = 4;  // J <= 128

__device__ __forceinline__ unsigned long long dkey(double c) {
  const long long b = __double_as_longlong(c);
  return static_cast<unsigned long long>(b) ^ (static_cast<unsigned long long>(b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
  return __longlong_as_double(static_cast<long long>(b));
}

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

// Shared-memory plan of one CTA: NS = kDpWarps * SEGS segment columns.
struct DpLayout {
  size_t rows, E, segend, Ej, trl, segb, fin_v, fin_j, bsS, bpS, total;
};
__host__ __device__ inline DpLayout dp_layout(int NS, int J, int bs_elem, int bp_rows) {
  DpLayout L;
  size_t o = 0;
  L.rows = o; o += sizeof(double) * (size_t)NS * J;
  L.E = o; o += sizeof(double) * 2 * NS;
  L.segend = o; o += sizeof(int64_t) * NS;
  L.fin_v = o; o += sizeof(double) * kDpWarps;
  L.Ej = o; o += sizeof(int) * 2 * NS;
  L.trl = o; o += sizeof(int) * NS;
  L.segb = o; o += sizeof(int) * NS;
  L.fin_j = o; o += sizeof(int) * kDpWarps;
  o = (o + 15) & ~size_t(15);
  L.bsS = o; o += (size_t)bs_elem * 2 * kDpChunk * NS;
  o = (o + 15) & ~size_t(15);
  L.bpS = o; o += (size_t)bp_rows * NS;
  L.total = (o + 15) & ~size_t(15);
  return L;
}

template <typename BST, int SLOTS, int SEGS>
__global__ void __launch_bounds__(kDpWarps * 32, (SEGS <= 2) ? 2 : 1)
dp_kernel(const mucon_viterbi_batch b, const int J, const int32_t* __restrict__ warp_unit, const int bp_rows) {
  extern __shared__ __align__(16) unsigned char sm[];
  constexpr int NS = kDpWarps * SEGS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* wu = warp_unit + static_cast<size_t>(blockIdx.x) * kDpWarps;
  const int u = wu[warp];
  if (u < 0) return;  // unused warp of a partially filled bin
  int w0 = warp, w1 = warp + 1;
  while (w0 > 0 && wu[w0 - 1] == u) --w0;
  while (w1 < kDpWarps && wu[w1] == u) ++w1;
  const int nw = w1 - w0, wl = warp - w0;
  const int ltid = wl * 32 + lane, nthr = nw * 32;
  const int s0 = w0 * SEGS;  // first segment column of this unit
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

  const DpLayout L = dp_layout(NS, J, sizeof(BST), bp_rows);
  double* rows = reinterpret_cast<double*>(sm + L.rows) + static_cast<size_t>(s0) * J;
  double* E = reinterpret_cast<double*>(sm + L.E);
  int64_t* segend = reinterpret_cast<int64_t*>(sm + L.segend) + s0;
  int* Ej = reinterpret_cast<int*>(sm + L.Ej);
  int* trl = reinterpret_cast<int*>(sm + L.trl) + s0;
  int* segb = reinterpret_cast<int*>(sm + L.segb) + s0;
  double* fin_v = reinterpret_cast<double*>(sm + L.fin_v) + w0;
  int* fin_j = reinterpret_cast<int*>(sm + L.fin_j) + w0;
  BST* bsS = reinterpret_cast<BST*>(sm + L.bsS);
  uint8_t* bpS = sm + L.bpS;
  const BST* bs_g = reinterpret_cast<const BST*>(b.bs) + b.blk_off[v] * C;
  uint8_t* bp_g = reinterpret_cast<uint8_t*>(b.bp) + b.bp_off[u];
  const bool bp_in_smem = bp_rows >= K;

  for (int n = ltid; n < N; n += nthr) trl[n] = b.tr[tr0 + n];
  // length rows: given, or ((l*ln m - m) - lf_l) - norms   (length_model.py:65-71,76-80)
  if (b.len_rows) {
    const double* g = b.len_rows + static_cast<size_t>(tr0) * J;
    for (int i = ltid; i < N * J; i += nthr) rows[i] = g[i];
  } else {
    const double* g = b.len_params + static_cast<size_t>(tr0) * 3;
    for (int i = ltid; i < N * J; i += nthr) {
      const int n = i / J, j = i - n * J + 1;
      const int l = j * fs;
      double r;
      if (l >= b.max_len) {
        r = -INFINITY;
      } else {
        r = __dmul_rn(static_cast<double>(l), g[n * 3 + 0]);
        r = __dsub_rn(r, g[n * 3 + 1]);
        r = __dsub_rn(r, b.logfact[j]);
        r = __dsub_rn(r, g[n * 3 + 2]);
      }
      rows[i] = r;
    }
  }
  ubar();  // trl visible

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
    auto stage = [&](int chunk) {
      const int k0 = chunk * kDpChunk;
      const int nk = min(kDpChunk, K - k0);
      BST* dst = bsS + static_cast<size_t>(chunk & 1) * kDpChunk * NS + s0;
      for (int i = ltid; i < nk * N; i += nthr) {
        const int kk = i / N, n = i - kk * N;
        const BST* src = bs_g + static_cast<int64_t>(k0 + kk) * C + trl[n];
        if (sizeof(BST) == 4) cp_async4(dst + kk * NS + n, src); else cp_async8(dst + kk * NS + n, src);
      }
      cp_async_commit();
    };
    const int nchunks = (K + kDpChunk - 1) / kDpChunk;
    stage(0);
    if (nchunks > 1) { stage(1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    ubar();

    double S[SEGS][SLOTS];
    int len[SEGS][SLOTS];
#pragma unroll
    for (int q = 0; q < SEGS; ++q)
#pragma unroll
      for (int i = 0; i < SLOTS; ++i) { S[q][i] = 0.0; len[q][i] = 0; }
    if (ltid == 0) {  // start hypothesis: 0.0 + F[fs-1, tr_0]  (viterbi.py:81-90)
      S[0][0] = __dadd_rn(0.0, static_cast<double>(bsS[s0]));
      len[0][0] = 1;
    }
    const bool f32seg0 = (sizeof(BST) == 4) && b.seg0_f32;

    int slot = (J > 1) ? 1 : 0;  // k mod J
    int kk = 1, chunk = 0;
    for (int k = 1; k < K; ++k) {
      if (kk == kDpChunk) {
        kk = 0;
        ++chunk;
        cp_async_wait<0>();
        ubar();  // chunk landed for every thread of the unit; the other buffer is free
        if (chunk + 1 < nchunks) stage(chunk + 1);
      }
      const BST* bsk = bsS + (static_cast<size_t>(chunk & 1) * kDpChunk + kk) * NS + s0;
      const int par = k & 1;
#pragma unroll
      for (int q = 0; q < SEGS; ++q) {
        const int n = wl * SEGS + q;
        if (n < N) {  // warp-uniform
          const BST bval = bsk[n];
          const double bd = static_cast<double>(bval);
          const double* row = rows + n * J - 1;
          unsigned long long key[SLOTS];
          unsigned long long best = 0;
#pragma unroll
          for (int i = 0; i < SLOTS; ++i) {
            const int ln = len[q][i];
            double a;
            if (f32seg0 && n == 0)
              a = static_cast<double>(__fadd_rn(static_cast<float>(S[q][i]), static_cast<float>(bval)));
            else
              a = __dadd_rn(S[q][i], bd);
            const double c = __dadd_rn(__dadd_rn(a, row[max(ln, 1)]), 0.0);
            key[i] = (ln > 0) ? dkey(c) : 0ull;
            best = (key[i] > best) ? key[i] : best;
            S[q][i] = a;
          }
          // warp max of the 64-bit key through two 32-bit REDUX
          const unsigned hi = static_cast<unsigned>(best >> 32);
          const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
          const unsigned lo = (hi == mh) ? static_cast<unsigned>(best) : 0u;
          const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
          const unsigned long long gk = (static_cast<unsigned long long>(mh) << 32) | ml;
          unsigned cl = 0;
#pragma unroll
          for (int i = 0; i < SLOTS; ++i) {
            const int ln = len[q][i];
            const unsigned cand = (key[i] == gk) ? static_cast<unsigned>(ln) : 0u;
            cl = max(cl, cand);
            len[q][i] = (ln > 0 && ln < J) ? ln + 1 : 0;
          }
          const unsigned jw = __reduce_max_sync(0xffffffffu, cl);  // 0 = no live predecessor
          if (lane == 0 && n + 1 < N) {
            E[par * NS + s0 + n + 1] = dunkey(gk);
            Ej[par * NS + s0 + n + 1] = static_cast<int>(jw);
            if (bp_in_smem) bpS[static_cast<size_t>(k) * NS + s0 + n + 1] = static_cast<uint8_t>(jw);
            else bp_g[static_cast<int64_t>(k) * N + n + 1] = static_cast<uint8_t>(jw);
          }
        }
      }
      ubar();
#pragma unroll
      for (int q = 0; q < SEGS; ++q) {
        const int n = wl * SEGS + q;
        if (n > 0 && n < N) {
          const int ej = Ej[par * NS + s0 + n];
          const double ev = E[par * NS + s0 + n];
#pragma unroll
          for (int i = 0; i < SLOTS; ++i)
            if (ej > 0 && lane + 32 * i == slot) { S[q][i] = ev; len[q][i] = 1; }
        }
      }
      ++kk;
      slot = (slot + 1 == J) ? 0 : slot + 1;
    }

    // end symbol: fold over the last segment (viterbi.py:125-138)
#pragma unroll
    for (int q = 0; q < SEGS; ++q) {
      const int n = wl * SEGS + q;
      if (n == N - 1) {
        const double* row = rows + n * J - 1;
        unsigned long long key[SLOTS];
        unsigned long long best = 0;
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) {
          const int ln = len[q][i];
          const double c = __dadd_rn(__dadd_rn(S[q][i], row[max(ln, 1)]), 0.0);
          key[i] = (ln > 0) ? dkey(c) : 0ull;
          best = (key[i] > best) ? key[i] : best;
        }
        const unsigned hi = static_cast<unsigned>(best >> 32);
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned lo = (hi == mh) ? static_cast<unsigned>(best) : 0u;
        const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
        const unsigned long long gk = (static_cast<unsigned long long>(mh) << 32) | ml;
        unsigned cl = 0;
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) cl = max(cl, (key[i] == gk) ? static_cast<unsigned>(len[q][i]) : 0u);
        const unsigned jw = __reduce_max_sync(0xffffffffu, cl);
        if (lane == 0) { *fin_v = dunkey(gk); *fin_j = static_cast<int>(jw); }
      }
    }
    ubar();
    if (ltid == 0) {  // traceback over the back-pointer table (viterbi.py:140-153)
      const double sc = *fin_v;
      const int jf = *fin_j;
      int n = N - 1;
      int k0 = K - jf;
      segb[n] = jf;
      while (n > 0) {
        const int ln = bp_in_smem ? static_cast<int>(bpS[static_cast<size_t>(k0) * NS + s0 + n])
                                  : static_cast<int>(__ldcg(bp_g + static_cast<int64_t>(k0) * N + n));
        segb[n - 1] = ln;
        k0 -= ln;
        --n;
      }
      b.score[u] = sc;
      b.final_j[u] = jf;
      b.status[u] = (isfinite(sc) || sc == -INFINITY) ? MUCON_UNIT_OK : MUCON_UNIT_NONFINITE;
    }
    ubar();
    // back-pointer table -> HBM, [K, N] row-major; row 0 and column 0 hold no entries
    if (bp_in_smem) {
      const int total = K * N;
      int k = ltid / N, n = ltid - k * N;
      const int dk = nthr / N, dn = nthr - dk * N;
      for (int i = ltid; i < total; i += nthr) {
        bp_g[i] = (k > 0 && n > 0) ? bpS[static_cast<size_t>(k) * NS + s0 + n] : uint8_t(0);
        k += dk;
        n += dn;
        if (n >= N) { n -= N; ++k; }
      }
    } else {
      for (int i = ltid; i < N; i += nthr) bp_g[i] = 0;
      for (int k = 1 + ltid; k < K; k += nthr) bp_g[static_cast<int64_t>(k) * N] = 0;
    }
  }

  for (int n = ltid; n < N; n += nthr) b.seg_blocks[tr0 + n] = segb[n];
  const int64_t lo = b.lab_off ? b.lab_off[u] : -1;
  if (lo >= 0) {
    if (ltid == 0) {
      int64_t pos = rem;
      for (int n = 0; n < N; ++n) { pos += static_cast<int64_t>(fs) * segb[n]; segend[n] = pos; }
    }
    ubar();
    write_labels(b.labels + lo, T, rem, trl, segend, last, ltid, nthr);
  }
}

}  // namespace mucon
