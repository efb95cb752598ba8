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
//                   (viterbi.py:92-138) with one warp per transcript segment, the J length
//                   slots of a segment spread over the lanes as a circular buffer indexed by
//                   entry step (no shifting), a warp-shuffle (value, length) arg-max with the
//                   reference's "last writer wins" tie rule (viterbi.py:26-28), back-pointers,
//                   traceback (viterbi.py:140-158) and a vectorised label writer.
//
// This translation unit is compiled with -fmad=false: every add/sub/mul must round exactly like
// the NumPy scalar arithmetic it replaces.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace mucon {
namespace {

constexpr int kMaxSlots = 4;    // J <= 128
constexpr int kDpMaxWarps = 16;  // CTA of at most 512 threads
constexpr int kDpChunk = 32;     // DP steps per block-score staging chunk

__host__ __device__ __forceinline__ int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

// ============================================================================================
// Block-score scan
// ============================================================================================

template <typename T>
__device__ __forceinline__ T neg_zero();
template <>
__device__ __forceinline__ float neg_zero<float>() { return -0.0f; }
template <>
__device__ __forceinline__ double neg_zero<double>() { return -0.0; }

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
// Dynamic program
// ============================================================================================

struct Best {
  double v;
  int j;  // 0 = no candidate
};

// max by value, ties -> larger j ("replace iff old <= new" over ascending j, viterbi.py:26-28).
__device__ __forceinline__ void best_take(Best& a, double v, int j) {
  const bool take = (j != 0) && ((a.j == 0) || (v > a.v) || (v == a.v && j > a.j));
  if (take) { a.v = v; a.j = j; }
}

__device__ __forceinline__ Best warp_best(Best a) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, a.v, off);
    const int oj = __shfl_xor_sync(0xffffffffu, a.j, off);
    best_take(a, ov, oj);
  }
  return a;
}

__device__ __forceinline__ int label_of_frame(int64_t t, int64_t rem, const int32_t* trl, const int64_t* segend,
                                              int last) {
  if (t < rem) return trl[last];
  int n = 0;
  while (n < last && t >= segend[n]) ++n;
  return trl[n];
}

// Writes T labels at out.  segend[n] = rem + fs * sum_{m<=n} blocks[m] (frames, exclusive end).
__device__ void write_labels(int32_t* out, int64_t T, int64_t rem, const int32_t* trl, const int64_t* segend,
                             int last) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int64_t mis = (reinterpret_cast<uintptr_t>(out) >> 2) & 3;
  const int64_t head = min64(T, (4 - mis) & 3);
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

struct DpSmem {
  double* rows;     // [N, J]
  double* E;        // [2, N] entry scores by step parity
  int64_t* segend;  // [N]
  int* Ej;          // [2, N] winning predecessor length (0 = no entry)
  int* trl;         // [N]
  int* segb;        // [N]
  void* bsS;        // [2, CH, N] staged block scores of the transcript labels
};

__host__ __device__ inline size_t dp_smem_bytes(int N, int J, int bs_elem) {
  size_t b = 0;
  b += sizeof(double) * (size_t)N * J;
  b += sizeof(double) * 2 * N;
  b += sizeof(int64_t) * N;
  b += sizeof(int) * 2 * N;
  b += sizeof(int) * N;
  b += sizeof(int) * N;
  b = (b + 15) & ~size_t(15);
  b += (size_t)bs_elem * 2 * kDpChunk * N;
  return b + 64;
}

template <typename BST, typename BPT, int SLOTS, int SEGS>
__global__ void __launch_bounds__(kDpMaxWarps * 32) dp_kernel(const mucon_viterbi_batch b, const int J) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ double fin_v;
  __shared__ int fin_j;

  const int u = b.order ? b.order[blockIdx.x] : blockIdx.x;
  const int v = b.unit_vid[u];
  const int64_t T = b.vid_off[v + 1] - b.vid_off[v];
  const int fs = b.fs;
  const int64_t K = T / fs;
  const int tr0 = b.tr_off[u];
  const int N = b.tr_off[u + 1] - tr0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = b.C;

  if (K < 1 || N < 1 || K > static_cast<int64_t>(N) * J) {
    if (tid == 0) {
      b.status[u] = MUCON_UNIT_INFEASIBLE;
      b.score[u] = __longlong_as_double(0x7ff8000000000000ll);
      b.final_j[u] = 0;
    }
    for (int n = tid; n < N; n += blockDim.x) b.seg_blocks[tr0 + n] = 0;
    return;
  }

  DpSmem s;
  {
    unsigned char* p = sm;
    s.rows = reinterpret_cast<double*>(p); p += sizeof(double) * (size_t)N * J;
    s.E = reinterpret_cast<double*>(p); p += sizeof(double) * 2 * N;
    s.segend = reinterpret_cast<int64_t*>(p); p += sizeof(int64_t) * N;
    s.Ej = reinterpret_cast<int*>(p); p += sizeof(int) * 2 * N;
    s.trl = reinterpret_cast<int*>(p); p += sizeof(int) * N;
    s.segb = reinterpret_cast<int*>(p); p += sizeof(int) * N;
    p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
    s.bsS = p;
  }
  BST* bsS = reinterpret_cast<BST*>(s.bsS);
  const BST* bs_g = reinterpret_cast<const BST*>(b.bs) + b.blk_off[v] * C;
  BPT* bp_g = b.bp ? reinterpret_cast<BPT*>(b.bp) + b.bp_off[u] : nullptr;

  for (int n = tid; n < N; n += blockDim.x) s.trl[n] = b.tr[tr0 + n];
  // length rows: given, or ((l*ln m - m) - lf_l) - norms   (length_model.py:65-71,76-80)
  if (b.len_rows) {
    const double* g = b.len_rows + static_cast<size_t>(tr0) * J;
    for (int i = tid; i < N * J; i += blockDim.x) s.rows[i] = g[i];
  } else {
    const double* g = b.len_params + static_cast<size_t>(tr0) * 3;
    for (int i = tid; i < N * J; i += blockDim.x) {
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
      s.rows[i] = r;
    }
  }
  if (bp_g) {  // row 0 has no entries
    for (int n = tid; n < N; n += blockDim.x) bp_g[n] = 0;
  }
  __syncthreads();  // trl visible

  const int Wa = (N + SEGS - 1) / SEGS;  // warps that own segments
  const int nact = Wa * 32;
  const int64_t rem = T - K * fs;
  int last = N - 1;

  if (K < N) {
    // Nothing reaches the last segment: the reference returns -inf and the path with one block
    // in each of the first K segments (viterbi.py:125-138).
    last = static_cast<int>(K) - 1;
    for (int n = tid; n < N; n += blockDim.x) s.segb[n] = (n < K) ? 1 : 0;
    if (tid == 0) {
      b.status[u] = MUCON_UNIT_SHORT;
      b.score[u] = -INFINITY;
      b.final_j[u] = 1;
    }
    if (bp_g) {
      for (int64_t i = N + tid; i < K * N; i += blockDim.x) bp_g[i] = 0;  // not computed
    }
    __syncthreads();
  } else {
    if (warp < Wa) {
      auto stage = [&](int chunk) {
        const int64_t k0 = static_cast<int64_t>(chunk) * kDpChunk;
        const int nk = static_cast<int>(min64(kDpChunk, K - k0));
        BST* dst = bsS + static_cast<size_t>(chunk & 1) * kDpChunk * N;
        for (int i = tid; i < nk * N; i += nact) {
          const int kk = i / N, n = i - kk * N;
          const BST* src = bs_g + (k0 + kk) * C + s.trl[n];
          if (sizeof(BST) == 4) cp_async4(dst + i, src); else cp_async8(dst + i, src);
        }
        cp_async_commit();
      };
      const int nchunks = static_cast<int>((K + kDpChunk - 1) / kDpChunk);
      stage(0);
      if (nchunks > 1) { stage(1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      named_bar_sync(1, nact);

      double S[SEGS][SLOTS];
      int len[SEGS][SLOTS];
#pragma unroll
      for (int q = 0; q < SEGS; ++q)
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) { S[q][i] = 0.0; len[q][i] = 0; }
      if (tid == 0) {  // start hypothesis: 0.0 + F[fs-1, tr_0]  (viterbi.py:81-90)
        S[0][0] = __dadd_rn(0.0, static_cast<double>(bsS[0]));
        len[0][0] = 1;
      }

      for (int64_t k = 1; k < K; ++k) {
        const int chunk = static_cast<int>(k / kDpChunk);
        const int kk = static_cast<int>(k - static_cast<int64_t>(chunk) * kDpChunk);
        if (kk == 0) {
          cp_async_wait<0>();
          named_bar_sync(1, nact);  // chunk landed for everyone; chunk-1 buffer is free
          if (chunk + 1 < nchunks) stage(chunk + 1);
        }
        const BST* bsk = bsS + (static_cast<size_t>(chunk & 1) * kDpChunk + kk) * N;
        const int par = static_cast<int>(k & 1);
#pragma unroll
        for (int q = 0; q < SEGS; ++q) {
          const int n = warp * SEGS + q;
          if (n < N) {
            const BST bval = bsk[n];
            const double* row = s.rows + n * J - 1;
            Best best{0.0, 0};
#pragma unroll
            for (int i = 0; i < SLOTS; ++i) {
              if (len[q][i] > 0) {
                double a;
                if (sizeof(BST) == 4 && b.seg0_f32 && n == 0)
                  a = static_cast<double>(__fadd_rn(static_cast<float>(S[q][i]), static_cast<float>(bval)));
                else
                  a = __dadd_rn(S[q][i], static_cast<double>(bval));
                const double cand = __dadd_rn(__dadd_rn(a, row[len[q][i]]), 0.0);
                best_take(best, cand, len[q][i]);
                S[q][i] = a;
                len[q][i] = (len[q][i] < J) ? len[q][i] + 1 : 0;
              }
            }
            best = warp_best(best);
            if (lane == 0) {
              if (n + 1 < N) {
                s.E[par * N + n + 1] = best.v;
                s.Ej[par * N + n + 1] = best.j;
                if (bp_g) bp_g[k * N + n + 1] = static_cast<BPT>(best.j);
              }
              if (n == 0 && bp_g) bp_g[k * N] = 0;
            }
          }
        }
        named_bar_sync(1, nact);
        const int slot = static_cast<int>(k % J);
#pragma unroll
        for (int q = 0; q < SEGS; ++q) {
          const int n = warp * SEGS + q;
          if (n > 0 && n < N) {
            const int ej = s.Ej[par * N + n];
            if (ej > 0) {
              const double ev = s.E[par * N + n];
#pragma unroll
              for (int i = 0; i < SLOTS; ++i)
                if (lane + 32 * i == slot) { S[q][i] = ev; len[q][i] = 1; }
            }
          }
        }
      }

      // end symbol: fold over the last segment (viterbi.py:125-138)
#pragma unroll
      for (int q = 0; q < SEGS; ++q) {
        const int n = warp * SEGS + q;
        if (n == N - 1) {
          const double* row = s.rows + n * J - 1;
          Best best{0.0, 0};
#pragma unroll
          for (int i = 0; i < SLOTS; ++i)
            if (len[q][i] > 0) best_take(best, __dadd_rn(__dadd_rn(S[q][i], row[len[q][i]]), 0.0), len[q][i]);
          best = warp_best(best);
          if (lane == 0) { fin_v = best.v; fin_j = best.j; }
        }
      }
    }
    __syncthreads();
    if (tid == 0) {  // traceback over the back-pointer table (viterbi.py:140-153)
      const double sc = fin_v;
      int n = N - 1;
      int64_t k0 = K - fin_j;
      s.segb[n] = fin_j;
      while (n > 0) {
        const int ln = static_cast<int>(__ldcg(bp_g + k0 * N + n));
        s.segb[n - 1] = ln;
        k0 -= ln;
        --n;
      }
      b.score[u] = sc;
      b.final_j[u] = fin_j;
      b.status[u] = (isfinite(sc) || sc == -INFINITY) ? MUCON_UNIT_OK : MUCON_UNIT_NONFINITE;
    }
    __syncthreads();
  }

  for (int n = tid; n < N; n += blockDim.x) b.seg_blocks[tr0 + n] = s.segb[n];
  const int64_t lo = b.lab_off ? b.lab_off[u] : -1;
  if (lo >= 0) {
    if (tid == 0) {
      int64_t pos = rem;
      for (int n = 0; n < N; ++n) { pos += static_cast<int64_t>(fs) * s.segb[n]; s.segend[n] = pos; }
    }
    __syncthreads();
    write_labels(b.labels + lo, T, rem, s.trl, s.segend, last);
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
  write_labels(labels + out_off[blockIdx.x], T, T - K * fs, trl, segend, last_s);
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
  const int slab_target = env_int("MUCON_SCAN_SLAB_BYTES", 6144);
  int bps = (int)(slab_target / blk_bytes);
  if (bps < 1) bps = 1;
  int stages = env_int("MUCON_SCAN_STAGES", 4);
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

template <typename BST, typename BPT, int SLOTS, int SEGS>
int launch_dp(const mucon_viterbi_batch& b, int J, cudaStream_t st) {
  const int Wa = (b.max_N + SEGS - 1) / SEGS;
  const size_t smem = dp_smem_bytes(b.max_N, J, sizeof(BST));
  auto kern = dp_kernel<BST, BPT, SLOTS, SEGS>;
  if (smem > 227 * 1024) return MUCON_EUNSUPPORTED;
  if (smem > 48 * 1024)
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<b.U, Wa * 32, smem, st>>>(b, J);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

template <typename BST, typename BPT, int SLOTS>
int dispatch_segs(const mucon_viterbi_batch& b, int J, cudaStream_t st) {
  const int segs = (b.max_N + kDpMaxWarps - 1) / kDpMaxWarps;
  if (segs <= 1) return launch_dp<BST, BPT, SLOTS, 1>(b, J, st);
  if (segs <= 2) return launch_dp<BST, BPT, SLOTS, 2>(b, J, st);
  if (segs <= 4) return launch_dp<BST, BPT, SLOTS, 4>(b, J, st);
  if (segs <= 8) return launch_dp<BST, BPT, SLOTS, 8>(b, J, st);
  return MUCON_EUNSUPPORTED;
}

template <typename BST, typename BPT>
int dispatch_slots(const mucon_viterbi_batch& b, int J, cudaStream_t st) {
  const int slots = (J + 31) / 32;
  switch (slots) {
    case 1: return dispatch_segs<BST, BPT, 1>(b, J, st);
    case 2: return dispatch_segs<BST, BPT, 2>(b, J, st);
    case 3: return dispatch_segs<BST, BPT, 3>(b, J, st);
    case 4: return dispatch_segs<BST, BPT, 4>(b, J, st);
    default: return MUCON_EUNSUPPORTED;
  }
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
  if (b.U < 0 || b.C < 1 || b.fs < 1 || b.max_len < b.fs || b.max_N < 1) return MUCON_EINVAL;
  if (!b.bs || !b.vid_off || !b.blk_off || !b.unit_vid || !b.tr || !b.tr_off || !b.score || !b.seg_blocks ||
      !b.final_j || !b.status || !b.bp || !b.bp_off)
    return MUCON_EINVAL;
  if (!b.len_rows && !(b.len_params && b.logfact)) return MUCON_EINVAL;
  if (b.U == 0) return MUCON_OK;
  const int J = b.max_len / b.fs;
  if (J > 32 * kMaxSlots) return MUCON_EUNSUPPORTED;
  if (!b.bp_is_u16 && J > 255) return MUCON_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // J <= 128 always fits uint8; the uint16 layout is reserved for the large-J path.
  if (b.bp_is_u16) return MUCON_EUNSUPPORTED;
  if (b.bs_is_f64) return dispatch_slots<double, uint8_t>(b, J, st);
  return dispatch_slots<float, uint8_t>(b, J, st);
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
