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
constexpr bool kStaged = true;   // hand-interleaved DP step (see dp_unit)
constexpr int kDpMaxJ = 128;     // ages live in registers: 4 lanes x <= 8 or 8 lanes x <= 16

// Lanes per segment.  A single warp issues roughly one instruction every 3.8 cycles on this
// dependent code (measured), so the per-step latency of a unit is set by the instructions per
// warp-step: 8 lanes x 9 ages for J = 66 keeps a step near 220 instructions while still packing
// four segments into a warp (4 lanes x 17 ages halves the warps but was measured slower overall).
// want == 32 gives every segment a warp of its own: the lane butterfly becomes three REDUX
// instructions and a step is ~70 instructions -- a third of the latency for three to four times the
// warps.  The host asks for it only for the few longest videos of a batch, whose serial chain of
// steps is the critical path (AlignPlan.n_long, mucon_viterbi_align_fused_tail).
__host__ __device__ inline int dp_group(int J, int max_N, int want) {
  const int shared = J <= 32 ? 4 : 8;  // segments sharing a warp
  if (want == 32) return (max_N <= 1 + kDpMaxWarps) ? 32 : shared;
  if (want == 4 || want == 8) return shared;
  return shared;
}
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

// --------------------------------------------------------------------------------------------
// One unit's thread team inside a CTA.
struct DpTeam {
  int u;          // unit index
  int lane, wl;   // lane, warp index within the team
  int nw;         // warps in the team
  int ltid, nthr; // thread index within the team, team size
  int bar_id;     // named barrier of the team (used when nw > 1)
  int slot;       // per-unit slot in the CTA's shared arrays (rows0, fin, Ex base)
  int c0;         // first segment column of the unit in the CTA's shared arrays
  int NS;         // segment columns of the CTA (row stride of bsS / bpS)
  __device__ __forceinline__ void sync() const {
    if (nw == 1) __syncwarp(); else named_bar_sync(bar_id, nthr);
  }
};

struct DpShared {
  double* rows0;    // [J] length scores of segment 0, ages 1..J
  double* Ex;       // [2][ex_stride] cross-warp entry scores by step parity, indexed by team warp
  int ex_stride;
  int64_t* segend;  // [N]
  int* trl;         // [N] transcript labels
  int* segb;        // [N] segment lengths in blocks
  double* fin_v;
  int* fin_j;
  uint8_t* bpS;     // [bp_rows][NS] back-pointer stage (column c0 + n)
  int bp_rows;
};

// ---- block-score sources --------------------------------------------------------------------
// Both sources hand out the block scores of one DP step at a time as two shared-memory loads (this
// lane's transcript label and the label of segment 0) from 32-bit shared addresses that advance by
// one row per step.
//
// StagedSrc: block scores come from HBM/L2 (written by the scan kernel); the team copies the
// columns of its transcript labels into a double-buffered shared stage, kDpChunk steps at a time.
template <typename BST>
struct StagedSrc {
  BST* bsS;          // [2][kDpChunk][NS]
  const BST* bs_g;   // video's [K][C] block scores
  const int* trl;
  int C, K, N, NS, c0, nchunks;
  int kk, chunk;     // position of the row last handed out
  uint32_t a_my, a_0, base_s;
  int off_my, off_0;
  __device__ __forceinline__ int col(int n, const int*) const { return n; }  // staged by segment
  __device__ __forceinline__ void bind(int my_col, int col0) {
    off_my = (c0 + my_col) * static_cast<int>(sizeof(BST));
    off_0 = (c0 + col0) * static_cast<int>(sizeof(BST));
  }
  __device__ __forceinline__ void stage(const DpTeam& t, int ch) {
    const int k0 = ch * kDpChunk;
    const int nk = min(kDpChunk, K - k0);
    BST* dst = bsS + static_cast<size_t>(ch & 1) * kDpChunk * NS + c0;
    for (int i = t.ltid; i < nk * N; i += t.nthr) {
      const int r = i / N, n = i - r * N;
      const BST* src = bs_g + static_cast<int64_t>(k0 + r) * C + trl[n];
      if (sizeof(BST) == 4) cp_async4(dst + r * NS + n, src); else cp_async8(dst + r * NS + n, src);
    }
    cp_async_commit();
  }
  // row 0 (returns after the first chunk has landed for the whole team)
  __device__ __forceinline__ void begin(const DpTeam& t) {
    nchunks = (K + kDpChunk - 1) / kDpChunk;
    stage(t, 0);
    if (nchunks > 1) { stage(t, 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    t.sync();
    kk = 0; chunk = 0;
    base_s = smem_u32(bsS);
    a_my = base_s + off_my;
    a_0 = base_s + off_0;
  }
  // row of the next step
  __device__ __forceinline__ void next(const DpTeam& t) {
    const uint32_t row_bytes = static_cast<uint32_t>(NS) * sizeof(BST);
    a_my += row_bytes;
    a_0 += row_bytes;
    if (++kk == kDpChunk) {
      kk = 0;
      ++chunk;
      cp_async_wait<0>();
      t.sync();  // chunk landed for every thread of the team; the other buffer is free
      if (chunk + 1 < nchunks) stage(t, chunk + 1);
      const uint32_t buf = base_s + static_cast<uint32_t>(chunk & 1) * kDpChunk * row_bytes;
      a_my = buf + off_my;
      a_0 = buf + off_0;
    }
  }
  __device__ __forceinline__ BST get_my() const { return ld_shared(a_my, BST()); }
  __device__ __forceinline__ BST get_0() const { return ld_shared(a_0, BST()); }
  __device__ __forceinline__ void finish(const DpTeam&) {}
};

// RingSrc: block scores are produced by the scan warps of the same CTA into a shared ring of
// slabs ([slab][bps][C], all classes); full/empty mbarriers hand slabs over.
template <typename BST>
struct RingSrc {
  const BST* ring;   // [slabs][bps][C]
  uint64_t* full;    // [slabs]
  uint64_t* empty;   // [slabs]
  int C, bps, slabs;
  int r, slab;       // row within slab, current slab
  uint32_t phase;    // parity of `full` for the current pass over the ring
  uint32_t a_my, a_0, ring_s, full_s, empty_s;
  int off_my, off_0;
  __device__ __forceinline__ int col(int n, const int* trl) const { return trl[n]; }  // by label
  __device__ __forceinline__ void bind(int my_col, int col0) {
    off_my = my_col * static_cast<int>(sizeof(BST));
    off_0 = col0 * static_cast<int>(sizeof(BST));
  }
  __device__ __forceinline__ void begin(const DpTeam&) {
    r = 0; slab = 0; phase = 0;
    ring_s = smem_u32(ring);
    full_s = smem_u32(full);
    empty_s = smem_u32(empty);
    a_my = ring_s + off_my;
    a_0 = ring_s + off_0;
    mbar_wait_s(full_s, 0);
  }
  __device__ __forceinline__ void next(const DpTeam& t) {
    const uint32_t row_bytes = static_cast<uint32_t>(C) * sizeof(BST);
    a_my += row_bytes;  // rows are contiguous across slabs; only the end of the ring wraps
    a_0 += row_bytes;
    if (++r == bps) {
      r = 0;
      // every thread of the team is past its reads of the old slab (the step ends with a team
      // barrier / warp sync), so one thread may hand it back
      t.sync();
      if (t.ltid == 0) mbar_arrive_s(empty_s + 8 * slab);
      if (++slab == slabs) {
        slab = 0;
        phase ^= 1;
        a_my = ring_s + off_my;
        a_0 = ring_s + off_0;
      }
      mbar_wait_s(full_s + 8 * slab, phase);
    }
  }
  __device__ __forceinline__ BST get_my() const { return ld_shared(a_my, BST()); }
  __device__ __forceinline__ BST get_0() const { return ld_shared(a_0, BST()); }
  __device__ __forceinline__ void finish(const DpTeam& t) {
    t.sync();
    if (t.ltid == 0) mbar_arrive_s(empty_s + 8 * slab);
  }
};

// --------------------------------------------------------------------------------------------
// The dynamic program, traceback and label writer of one unit.
template <typename BST, int G, int SL, typename Src>
__device__ __forceinline__ void dp_unit(const mucon_viterbi_batch& b, const int J, const DpTeam& t,
                                        const DpShared& sh, Src& src) {
  constexpr int kSegsPerWarp = 32 / G;
  const int u = t.u, lane = t.lane, wl = t.wl, nw = t.nw, ltid = t.ltid, nthr = t.nthr;
  const int v = b.unit_vid[u];
  const int64_t T = b.vid_off[v + 1] - b.vid_off[v];
  const int fs = b.fs;
  const int K = static_cast<int>(T / fs);
  const int tr0 = b.tr_off[u];
  const int N = b.tr_off[u + 1] - tr0;
  double* rows0 = sh.rows0;
  int64_t* segend = sh.segend;
  int* trl = sh.trl;
  int* segb = sh.segb;
  uint8_t* bp_g = b.bp + b.bp_off[u];
  const bool bp_in_smem = sh.bp_rows >= K;
  const int64_t rem = T - static_cast<int64_t>(K) * fs;
  int last = N - 1;

  for (int j = ltid; j < J; j += nthr) rows0[j] = length_row(b, tr0, 0, j + 1, J);

  if (K < N) {
    // Nothing reaches the last segment: the reference returns -inf and the path with one block
    // in each of the first K segments (viterbi.py:125-138).
    last = K - 1;
    for (int n = ltid; n < N; n += nthr) segb[n] = (n < K) ? 1 : 0;
    if (ltid == 0) {
      b.status[u] = MUCON_UNIT_SHORT;
      put_score(b, u, -INFINITY);
      b.final_j[u] = 1;
    }
    for (int i = ltid; i < K * N; i += nthr) bp_g[i] = 0;  // not computed
    // drain the source so that a producer never waits for this team
    src.bind(0, 0);
    src.begin(t);
    for (int k = 1; k < K; ++k) src.next(t);
    src.finish(t);
    t.sync();
  } else {
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
    const int my_col = src.col(has_seg ? n : 0, trl);
    const int col0 = src.col(0, trl);

    src.bind(my_col, col0);
    src.begin(t);  // block scores of step 0
    // segment 0: scalar chain; 0.0 + F[fs-1, tr_0]  (viterbi.py:81-90)
    const bool f32seg0 = (sizeof(BST) == 4) && b.seg0_f32;
    double s0 = __dadd_rn(0.0, static_cast<double>(src.get_0()));

    const unsigned gmask = (G == 32) ? 0xffffffffu : ((0xffffffffu >> (32 - G)) << (g * G));
    bool tie_ok[5];  // butterfly level lv: the partner is the higher lane (older ages), ties go to it
#pragma unroll
    for (int lv = 0; lv < 5; ++lv) tie_ok[lv] = (lane & (1 << lv)) == 0;
    const bool bp_writer = has_seg && lig == G - 1 && n + 1 < N;
    uint8_t* bp_w;  // where this lane records the winner of its group at step k
    int bp_stride;
    if (bp_in_smem) { bp_w = sh.bpS + t.c0 + n + 1; bp_stride = t.NS; }
    else { bp_w = bp_g + n + 1; bp_stride = N; }
    bp_w += bp_stride;  // step 1
    double* ex_out = sh.Ex + wl + 1;
    const double* ex_in = sh.Ex + wl;
    const bool ex_writer = lane == 31 && wl + 1 < nw;
    const bool ex_reader = lane == 0 && wl > 0;
    const bool multi = nw > 1;

    // The step is software-pipelined.  Only the youngest hypothesis of a segment depends on the
    // previous step's fold (it IS that fold's winner); everything else is a pure shift-and-add.
    //   A'(k): out_k, R[i>=2], candidates and tree fold of positions >= 1  -- no dependence on
    //          the entry produced by step k-1
    //   C(k):  R[1] = R[0] + b_k, youngest candidate, lane butterfly / REDUX, hand-over to the
    //          next segment / warp                              -- the loop-carried chain
    // The loop body is C(k); A'(k+1), so the shuffle latencies of C(k) are filled with the
    // independent arithmetic of A'(k+1).
    double out = 0.0, tv = -INFINITY, e1 = -INFINITY, bd = 0.0;
    int ti = 1;
    auto a_prime = [&](int k, double bdn, BST b0n, double& outn, double& tvn, int& tin, double& e1n) {
      if (SL > 1) {
        outn = __dadd_rn(R[SL - 1], bdn);
#pragma unroll
        for (int i = SL - 1; i >= 2; --i) R[i] = __dadd_rn(R[i - 1], bdn);
        double cv[SL];
        int idx[SL];
#pragma unroll
        for (int i = 1; i < SL; ++i) {
          const double a = (i + 1 < SL) ? R[(i + 1 < SL) ? i + 1 : 0] : outn;
          cv[i] = __dadd_rn(a, rowr[i]);
          idx[i] = i;
        }
#pragma unroll
        for (int w = 1; w < SL - 1; w <<= 1) {
#pragma unroll
          for (int i = 1; i + w < SL; i += 2 * w) {
            const bool older = cv[i + w] >= cv[i];  // the older position wins ties
            cv[i] = older ? cv[i + w] : cv[i];
            idx[i] = older ? idx[i + w] : idx[i];
          }
        }
        tvn = cv[(SL > 1) ? 1 : 0];
        tin = idx[(SL > 1) ? 1 : 0];
      }
      // segment 0 -> entry of segment 1.  Every warp runs the chain (four branch-free
      // instructions); only lane 0 of the team's first warp uses the result.
      double a;
      if (f32seg0) a = static_cast<double>(__fadd_rn(static_cast<float>(s0), static_cast<float>(b0n)));
      else a = __dadd_rn(s0, static_cast<double>(b0n));
      s0 = a;
      // the single hypothesis of segment 0 has age k; it can advance while k <= J.
      // rows are never -0.0, so the reference's trailing "+ 0.0" is the identity here.
      e1n = (k <= J) ? __dadd_rn(a, rows0[min(k, J) - 1]) : -INFINITY;
    };
    auto c_step = [&](int k, double& bv, double& inc) {
      int bi;
      // the shift between lanes does not depend on this step's fold: issue it first, so that the
      // register holding `out` is dead before A'(k+1) and its shift-adds can fill the butterfly's
      // latency (a single shuffle after the butterfly serialised the two phases through a WAR hazard)
      double inc_shift = 0.0;
      if (SL > 1) inc_shift = __shfl_up_sync(0xffffffffu, out, 1);
      if (SL > 1) {
        R[(SL > 1) ? 1 : 0] = __dadd_rn(R[0], bd);
        const double c0v = __dadd_rn(R[(SL > 1) ? 1 : 0], rowr[0]);
        const bool young = c0v > tv;  // the youngest wins only when strictly greater
        bv = young ? c0v : tv;
        bi = young ? 0 : ti;
      } else {
        out = __dadd_rn(R[0], bd);
        bv = __dadd_rn(out, rowr[0]);
        bi = 0;
      }
      int bage = a0 + bi + 1;
      if (G == 32) {
        // whole-warp arg-max on the order-preserving integer key: two REDUX for the 64-bit
        // maximum, one for the oldest age among the lanes that hold it.  (REDUX with per-group
        // member masks is serialised by the compiler, so smaller groups use the butterfly.)
        const unsigned long long key = dkey(bv);
        const unsigned hi = static_cast<unsigned>(key >> 32);
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned lo = (hi == mh) ? static_cast<unsigned>(key) : 0u;
        const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
        const bool mine = (hi == mh) && (static_cast<unsigned>(key) == ml);
        bage = static_cast<int>(__reduce_max_sync(0xffffffffu, mine ? static_cast<unsigned>(bage) : 0u));
        bv = dunkey((static_cast<unsigned long long>(mh) << 32) | ml);
      } else {
        // butterfly over the group's lanes: the partner with the higher lane holds older ages,
        // so it wins ties; the partner with the lower lane must be strictly greater
#pragma unroll
        for (int lv = 0; (1 << lv) < G; ++lv) {
          const double ov = __shfl_xor_sync(0xffffffffu, bv, 1 << lv);
          const int oa = __shfl_xor_sync(0xffffffffu, bage, 1 << lv);
          const bool take = (ov > bv) || (tie_ok[lv] && ov == bv);
          bv = take ? ov : bv;
          bage = take ? oa : bage;
        }
      }
      // shift between lanes; the first lane of a group takes the previous group's winner instead
      if (SL > 1) {
        if (G < 32) {
          const double win = __shfl_up_sync(0xffffffffu, bv, 1);
          inc = (lig == 0) ? win : inc_shift;
        } else {
          inc = inc_shift;
        }
      } else {
        const double send = (lig == G - 1) ? bv : out;
        inc = __shfl_up_sync(0xffffffffu, send, 1);
      }
      inc = (lane == 0) ? e1 : inc;  // first warp: from segment 0; other warps: patched later
      // a fold whose maximum is -inf is decided by liveness alone: oldest live age or none
      const int jhi = min(J, k - n), jlo = max(1, k - nJ);
      const int dead_age = (jlo <= jhi) ? jhi : 0;
      bage = (bv == -INFINITY) ? dead_age : bage;
      if (bp_writer) *bp_w = static_cast<uint8_t>(bage);
      bp_w += bp_stride;
    };
    auto exchange = [&](int k, double bv, double& inc) {
      if (multi) {
        const int par = (k & 1) * sh.ex_stride;
        if (ex_writer) ex_out[par] = bv;
        named_bar_sync(t.bar_id, nthr);
        if (ex_reader) inc = ex_in[par];
      }
    };

    if (K > 1) {
      src.next(t);
      bd = static_cast<double>(src.get_my());
      a_prime(1, bd, src.get_0(), out, tv, ti, e1);
    }
    for (int k = 1; k + 1 < K; ++k) {
      src.next(t);  // block scores of step k+1
      const double bdn = static_cast<double>(src.get_my());
      const BST b0n = src.get_0();
      double bv, inc, outn = 0.0, tvn = -INFINITY, e1n;
      int tin = 1;
      if (kStaged && G == 8 && SL > 2) {
        // C(k) and A'(k+1) interleaved by hand.  A warp issues in order and ptxas schedules the
        // butterfly (the critical path of the block) first and all of A' after it, so every
        // shuffle -> compare -> select hop stalls the warp with nothing to issue.  The warp-level
        // fences below end the scheduling region after each shuffle has been issued together with
        // a slice of A'(k+1): the slice runs while the shuffle is in flight.
        const double inc_shift = __shfl_up_sync(0xffffffffu, out, 1);
        R[1] = __dadd_rn(R[0], bd);
        const double c0v = __dadd_rn(R[1], rowr[0]);
        const bool young = c0v > tv;
        bv = young ? c0v : tv;
        int bage = a0 + (young ? 0 : ti) + 1;
        double ov = __shfl_xor_sync(0xffffffffu, bv, 1);
        int oa = __shfl_xor_sync(0xffffffffu, bage, 1);
        // slice 0: ageing
        outn = __dadd_rn(R[SL - 1], bdn);
#pragma unroll
        for (int i = SL - 1; i >= 2; --i) R[i] = __dadd_rn(R[i - 1], bdn);
        __syncwarp();
        bool take = (ov > bv) || (tie_ok[0] && ov == bv);
        bv = take ? ov : bv;
        bage = take ? oa : bage;
        ov = __shfl_xor_sync(0xffffffffu, bv, 2);
        oa = __shfl_xor_sync(0xffffffffu, bage, 2);
        // slice 1: candidates of positions >= 1 and the first level of their fold
        double cv[SL];
        int idx[SL];
#pragma unroll
        for (int i = 1; i < SL; ++i) {
          const double a = (i + 1 < SL) ? R[(i + 1 < SL) ? i + 1 : 0] : outn;
          cv[i] = __dadd_rn(a, rowr[i]);
          idx[i] = i;
        }
#pragma unroll
        for (int i = 1; i + 1 < SL; i += 2) {
          const bool older = cv[i + 1] >= cv[i];
          cv[i] = older ? cv[i + 1] : cv[i];
          idx[i] = older ? idx[i + 1] : idx[i];
        }
        __syncwarp();
        take = (ov > bv) || (tie_ok[1] && ov == bv);
        bv = take ? ov : bv;
        bage = take ? oa : bage;
        ov = __shfl_xor_sync(0xffffffffu, bv, 4);
        oa = __shfl_xor_sync(0xffffffffu, bage, 4);
        // slice 2: the rest of the fold
#pragma unroll
        for (int w = 2; w < SL - 1; w <<= 1) {
#pragma unroll
          for (int i = 1; i + w < SL; i += 2 * w) {
            const bool older = cv[i + w] >= cv[i];
            cv[i] = older ? cv[i + w] : cv[i];
            idx[i] = older ? idx[i + w] : idx[i];
          }
        }
        tvn = cv[1];
        tin = idx[1];
        __syncwarp();
        take = (ov > bv) || (tie_ok[2] && ov == bv);
        bv = take ? ov : bv;
        bage = take ? oa : bage;
        const double win = __shfl_up_sync(0xffffffffu, bv, 1);
        // slice 3: segment 0, liveness, back-pointer
        {
          double a;
          if (f32seg0) a = static_cast<double>(__fadd_rn(static_cast<float>(s0), static_cast<float>(b0n)));
          else a = __dadd_rn(s0, static_cast<double>(b0n));
          s0 = a;
          e1n = (k + 1 <= J) ? __dadd_rn(a, rows0[min(k + 1, J) - 1]) : -INFINITY;
        }
        const int jhi = min(J, k - n), jlo = max(1, k - nJ);
        const int dead_age = (jlo <= jhi) ? jhi : 0;
        bage = (bv == -INFINITY) ? dead_age : bage;
        if (bp_writer) *bp_w = static_cast<uint8_t>(bage);
        bp_w += bp_stride;
        __syncwarp();
        inc = (lig == 0) ? win : inc_shift;
        inc = (lane == 0) ? e1 : inc;
      } else {
        // C(k) and A'(k+1) in one basic block
        c_step(k, bv, inc);
        a_prime(k + 1, bdn, b0n, outn, tvn, tin, e1n);
      }
      exchange(k, bv, inc);
      R[0] = inc;
      out = outn; tv = tvn; ti = tin; e1 = e1n; bd = bdn;
    }
    if (K > 1) {  // last step: C(K-1) only
      double bv, inc;
      c_step(K - 1, bv, inc);
      exchange(K - 1, bv, inc);
      R[0] = inc;
    }
    src.finish(t);

    // end symbol: fold over the last segment (viterbi.py:125-138)
    if (N == 1) {
      if (ltid == 0) {
        *sh.fin_v = __dadd_rn(__dadd_rn(s0, rows0[K - 1]), 0.0);  // K <= J is guaranteed by feasibility
        *sh.fin_j = K;
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
        const double ov = __shfl_xor_sync(gmask, bv, off);
        const int oa = __shfl_xor_sync(gmask, bage, off);
        const bool take = (lane & off) ? (ov > bv) : (ov >= bv);
        if (take) { bv = ov; bage = oa; }
      }
      bv = __dadd_rn(bv, 0.0);
      // after the last step (k = K-1) the live ages of segment n are [max(1, K-n*J), min(J, K-n)]
      const int jhi = min(J, K - n), jlo = max(1, K - nJ);
      if (bv == -INFINITY) bage = (jlo <= jhi) ? jhi : 0;
      if (lig == 0) { *sh.fin_v = bv; *sh.fin_j = bage; }
    }
    t.sync();
    if (ltid == 0) {  // traceback over the back-pointer table (viterbi.py:140-153)
      const double sc = *sh.fin_v;
      const int jf = *sh.fin_j;
      int m = N - 1;
      int k0 = K - jf;
      segb[m] = jf;
      while (m > 0) {
        int ln;  // column 1 (entries from segment 0) is a function of the step alone
        if (m == 1) ln = (k0 <= J) ? k0 : 0;
        else ln = bp_in_smem ? static_cast<int>(sh.bpS[static_cast<size_t>(k0) * t.NS + t.c0 + m])
                             : static_cast<int>(__ldcg(bp_g + static_cast<int64_t>(k0) * N + m));
        segb[m - 1] = ln;
        k0 -= ln;
        --m;
      }
      put_score(b, u, sc);
      b.final_j[u] = jf;
      b.status[u] = (isfinite(sc) || sc == -INFINITY) ? MUCON_UNIT_OK : MUCON_UNIT_NONFINITE;
    }
    t.sync();
    // back-pointer table -> HBM, [K, N] row-major; row 0 and column 0 hold no entries
    if (bp_in_smem) {
      const int total = K * N;
      int k = ltid / N, m = ltid - k * N;
      const int dk = nthr / N, dn = nthr - dk * N;
      for (int i = ltid; i < total; i += nthr) {
        uint8_t val = 0;
        if (k > 0 && m == 1) val = (k <= J) ? static_cast<uint8_t>(k) : uint8_t(0);
        else if (k > 0 && m > 1) val = sh.bpS[static_cast<size_t>(k) * t.NS + t.c0 + m];
        bp_g[i] = val;
        k += dk;
        m += dn;
        if (m >= N) { m -= N; ++k; }
      }
    } else {
      for (int i = ltid; i < N; i += nthr) bp_g[i] = 0;
      for (int k = 1 + ltid; k < K; k += nthr) {
        bp_g[static_cast<int64_t>(k) * N] = 0;
        if (N > 1) bp_g[static_cast<int64_t>(k) * N + 1] = (k <= J) ? static_cast<uint8_t>(k) : uint8_t(0);
      }
    }
  }

  for (int m = ltid; m < N; m += nthr) put_seg(b, tr0 + m, segb[m]);
  const int64_t lo = b.lab_off ? b.lab_off[u] : -1;
  if (lo >= 0) {
    if (ltid == 0) {
      int64_t pos = rem;
      for (int m = 0; m < N; ++m) { pos += static_cast<int64_t>(fs) * segb[m]; segend[m] = pos; }
    }
    t.sync();
    write_labels(b.labels + lo, T, rem, trl, segend, last, ltid, nthr);
  }
}

// Unit-level feasibility (uniform over the team).  Returns false after writing the status.
__device__ __forceinline__ bool dp_feasible(const mucon_viterbi_batch& b, int J, const DpTeam& t) {
  const int u = t.u;
  const int v = b.unit_vid[u];
  const int64_t T = b.vid_off[v + 1] - b.vid_off[v];
  const int64_t K = T / b.fs;
  const int tr0 = b.tr_off[u];
  const int N = b.tr_off[u + 1] - tr0;
  if (K >= 1 && N >= 1 && K <= static_cast<int64_t>(N) * J) return true;
  if (t.ltid == 0) {
    b.status[u] = MUCON_UNIT_INFEASIBLE;
    put_score(b, u, __longlong_as_double(0x7ff8000000000000ll));
    b.final_j[u] = 0;
  }
  for (int n = t.ltid; n < N; n += t.nthr) put_seg(b, tr0 + n, 0);
  return false;
}

// --------------------------------------------------------------------------------------------
// Kernel 1: DP over block scores already in HBM (any number of candidates per video).
template <typename BST, int G, int SL>
__global__ void __launch_bounds__(kDpMaxWarps * 32, 1)
dp_kernel(const mucon_viterbi_batch b, const int J, const int32_t* __restrict__ warp_unit, const int bp_rows) {
  extern __shared__ __align__(16) unsigned char sm[];
  constexpr int kSegsPerWarp = 32 / G;
  const int wpc = blockDim.x >> 5, NS = dp_columns(wpc, G);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* wu = warp_unit + static_cast<size_t>(blockIdx.x) * wpc;
  const int u = wu[warp];
  if (u < 0) return;  // unused warp of a partially filled bin
  int w0 = warp, w1 = warp + 1;
  while (w0 > 0 && wu[w0 - 1] == u) --w0;
  while (w1 < wpc && wu[w1] == u) ++w1;
  DpTeam t;
  t.u = u; t.lane = lane; t.wl = warp - w0; t.nw = w1 - w0;
  t.ltid = t.wl * 32 + lane; t.nthr = t.nw * 32;
  t.bar_id = 1 + w0; t.slot = w0; t.c0 = w0 * (kSegsPerWarp + 1); t.NS = NS;
  if (!dp_feasible(b, J, t)) return;

  const DpLayout L = dp_layout(wpc, G, J, sizeof(BST), bp_rows);
  DpShared sh;
  sh.rows0 = reinterpret_cast<double*>(sm + L.rows0) + static_cast<size_t>(w0) * J;
  sh.Ex = reinterpret_cast<double*>(sm + L.Ex) + w0;
  sh.ex_stride = wpc;
  sh.segend = reinterpret_cast<int64_t*>(sm + L.segend) + t.c0;
  sh.trl = reinterpret_cast<int*>(sm + L.trl) + t.c0;
  sh.segb = reinterpret_cast<int*>(sm + L.segb) + t.c0;
  sh.fin_v = reinterpret_cast<double*>(sm + L.fin_v) + w0;
  sh.fin_j = reinterpret_cast<int*>(sm + L.fin_j) + w0;
  sh.bpS = sm + L.bpS;
  sh.bp_rows = bp_rows;

  const int v = b.unit_vid[u];
  const int tr0 = b.tr_off[u];
  const int N = b.tr_off[u + 1] - tr0;
  for (int n = t.ltid; n < N; n += t.nthr) sh.trl[n] = b.tr[tr0 + n];
  t.sync();  // trl visible

  StagedSrc<BST> src;
  src.bsS = reinterpret_cast<BST*>(sm + L.bsS);
  src.bs_g = reinterpret_cast<const BST*>(b.bs) + b.blk_off[v] * b.C;
  src.trl = sh.trl;
  src.C = b.C;
  src.K = static_cast<int>((b.vid_off[v + 1] - b.vid_off[v]) / b.fs);
  src.N = N; src.NS = NS; src.c0 = t.c0;
  dp_unit<BST, G, SL>(b, J, t, sh, src);
}

}  // namespace mucon
