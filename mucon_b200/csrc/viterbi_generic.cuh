// viterbi_generic.cuh -- the same dynamic program for shapes the register-resident kernels do not
// cover: J = max_len / fs > 128 (e.g. the reference Viterbi's default frame_sampling = 1 with
// max_length = 2000) or very long transcripts.  One CTA per unit, one warp per transcript segment
// (looping when N exceeds the warps), hypothesis scores in a global-memory workspace laid out as a
// circular buffer per segment (slot = entry step mod J, so nothing is ever shifted), length scores
// precomputed next to it.  Same arithmetic, tie rule and outputs as dp_unit (viterbi_dp.cuh);
// throughput is not a goal here, exactness is.
#pragma once
#include <math.h>

#include "viterbi_dp.cuh"

namespace mucon {

constexpr int kGenThreads = 512;
__device__ __forceinline__ int64_t gmin64(int64_t a, int64_t b) { return a < b ? a : b; }

template <typename BST, typename BPT>
__global__ void __launch_bounds__(kGenThreads)
dp_generic_kernel(const mucon_viterbi_batch b, const int J, double* __restrict__ ws,
                  const int64_t* __restrict__ ws_off) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ double fin_v;
  __shared__ int fin_j;
  const int u = blockIdx.x;
  const int v = b.unit_vid[u];
  const int64_t T = b.vid_off[v + 1] - b.vid_off[v];
  const int fs = b.fs;
  const int64_t K = T / fs;
  const int tr0 = b.tr_off[u];
  const int N = b.tr_off[u + 1] - tr0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int C = b.C;

  if (K < 1 || N < 1 || K > static_cast<int64_t>(N) * J) {
    if (tid == 0) {
      b.status[u] = MUCON_UNIT_INFEASIBLE;
      put_score(b, u, __longlong_as_double(0x7ff8000000000000ll));
      b.final_j[u] = 0;
    }
    for (int n = tid; n < N; n += blockDim.x) put_seg(b, tr0 + n, 0);
    return;
  }
  // shared: E[N] entry scores, Ej[N], trl[N], segb[N], segend[N]
  double* E = reinterpret_cast<double*>(sm);
  int64_t* segend = reinterpret_cast<int64_t*>(E + N);
  int* Ej = reinterpret_cast<int*>(segend + N);
  int* trl = Ej + N;
  int* segb = trl + N;
  double* S = ws + ws_off[u];                     // [N][J] scores by slot
  double* rows = S + static_cast<size_t>(N) * J;  // [N][J] length scores by age-1
  BPT* bp_g = reinterpret_cast<BPT*>(b.bp) + b.bp_off[u];
  const BST* bs_g = reinterpret_cast<const BST*>(b.bs) + b.blk_off[v] * C;
  const int64_t rem = T - K * fs;
  int last = N - 1;

  for (int n = tid; n < N; n += blockDim.x) trl[n] = b.tr[tr0 + n];
  for (int64_t i = tid; i < static_cast<int64_t>(N) * J; i += blockDim.x) {
    const int n = static_cast<int>(i / J), j = static_cast<int>(i - static_cast<int64_t>(n) * J) + 1;
    rows[i] = length_row(b, tr0, n, j, J);
    S[i] = -INFINITY;
  }
  for (int64_t i = tid; i < K * N; i += blockDim.x) bp_g[i] = 0;
  __syncthreads();

  if (K < N) {
    last = static_cast<int>(K) - 1;
    for (int n = tid; n < N; n += blockDim.x) segb[n] = (n < K) ? 1 : 0;
    if (tid == 0) {
      b.status[u] = MUCON_UNIT_SHORT;
      put_score(b, u, -INFINITY);
      b.final_j[u] = 1;
    }
    __syncthreads();
  } else {
    const bool f32seg0 = (sizeof(BST) == 4) && b.seg0_f32;
    if (tid == 0) S[0] = __dadd_rn(0.0, static_cast<double>(bs_g[trl[0]]));  // segment 0, slot 0 (entered at step 0)
    __syncthreads();
    for (int64_t k = 1; k < K; ++k) {
      const int slot_now = static_cast<int>(k % J);  // slot that receives this step's entry
      const int base = static_cast<int>((k - 1) % J);
      for (int n = warp; n < N; n += nwarp) {
        const BST bval = bs_g[k * C + trl[n]];
        double* Sn = S + static_cast<size_t>(n) * J;
        const double* rn = rows + static_cast<size_t>(n) * J;
        double bv = -INFINITY;
        int bage = 0;
        for (int s = lane; s < J; s += 32) {
          // hypothesis in slot s entered at the latest step k0 <= k-1 with k0 = s (mod J); age = k - k0
          int age = base - s;
          age += (age < 0) ? J + 1 : 1;
          double a;
          if (f32seg0 && n == 0) a = static_cast<double>(__fadd_rn(static_cast<float>(Sn[s]), static_cast<float>(bval)));
          else a = __dadd_rn(Sn[s], static_cast<double>(bval));
          const double c = __dadd_rn(a, rn[age - 1]);
          if (c > bv || (c == bv && age > bage) || bage == 0) { bv = c; bage = age; }
          Sn[s] = (age < J) ? a : -INFINITY;  // a hypothesis of J blocks cannot stay (viterbi.py:97)
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
          const int oa = __shfl_xor_sync(0xffffffffu, bage, off);
          if (oa != 0 && (bage == 0 || ov > bv || (ov == bv && oa > bage))) { bv = ov; bage = oa; }
        }
        if (lane == 0 && n + 1 < N) {
          // a fold whose maximum is -inf is decided by liveness alone (see dp_unit)
          int jhi, jlo;
          if (n == 0) { jhi = (k <= J) ? static_cast<int>(k) : 0; jlo = jhi > 0 ? jhi : 1; }
          else {
            jhi = static_cast<int>(gmin64(J, k - n));
            jlo = static_cast<int>(k - static_cast<int64_t>(n) * J > 1 ? k - static_cast<int64_t>(n) * J : 1);
          }
          if (bv == -INFINITY || bage == 0) bage = (jhi >= jlo && jhi >= 1) ? jhi : 0;
          E[n + 1] = __dadd_rn(bv, 0.0);
          Ej[n + 1] = bage;
          bp_g[k * N + n + 1] = static_cast<BPT>(bage);
        }
      }
      __syncthreads();
      for (int n = 1 + tid; n < N; n += blockDim.x) S[static_cast<size_t>(n) * J + slot_now] = (Ej[n] > 0) ? E[n] : -INFINITY;
      __syncthreads();
    }
    // end symbol: fold over the last segment
    if (warp == 0) {
      const int n = N - 1;
      const double* Sn = S + static_cast<size_t>(n) * J;
      const double* rn = rows + static_cast<size_t>(n) * J;
      double bv = -INFINITY;
      int bage = 0;
      const int base = static_cast<int>((K - 1) % J);
      for (int s = lane; s < J; s += 32) {
        int age = base - s;
        age += (age < 0) ? J + 1 : 1;
        const double c = __dadd_rn(Sn[s], rn[age - 1]);
        if (c > bv || (c == bv && age > bage) || bage == 0) { bv = c; bage = age; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oa = __shfl_xor_sync(0xffffffffu, bage, off);
        if (oa != 0 && (bage == 0 || ov > bv || (ov == bv && oa > bage))) { bv = ov; bage = oa; }
      }
      if (lane == 0) {
        int jhi, jlo;
        if (n == 0) { jhi = static_cast<int>(K); jlo = jhi; }
        else {
          jhi = static_cast<int>(gmin64(J, K - n));
          jlo = static_cast<int>(K - static_cast<int64_t>(n) * J > 1 ? K - static_cast<int64_t>(n) * J : 1);
        }
        if (bv == -INFINITY) bage = (jhi >= jlo) ? jhi : 0;
        fin_v = __dadd_rn(bv, 0.0);
        fin_j = bage;
      }
    }
    __syncthreads();
    if (tid == 0) {
      const double sc = fin_v;
      int n = N - 1;
      int64_t k0 = K - fin_j;
      segb[n] = fin_j;
      while (n > 0) {
        const int ln = static_cast<int>(__ldcg(bp_g + k0 * N + n));
        segb[n - 1] = ln;
        k0 -= ln;
        --n;
      }
      put_score(b, u, sc);
      b.final_j[u] = fin_j;
      b.status[u] = (isfinite(sc) || sc == -INFINITY) ? MUCON_UNIT_OK : MUCON_UNIT_NONFINITE;
    }
    __syncthreads();
  }
  for (int n = tid; n < N; n += blockDim.x) put_seg(b, tr0 + n, segb[n]);
  const int64_t lo = b.lab_off ? b.lab_off[u] : -1;
  if (lo >= 0) {
    if (tid == 0) {
      int64_t pos = rem;
      for (int n = 0; n < N; ++n) { pos += static_cast<int64_t>(fs) * segb[n]; segend[n] = pos; }
    }
    __syncthreads();
    write_labels(b.labels + lo, T, rem, trl, segend, last, tid, blockDim.x);
  }
}

}  // namespace mucon
