// metrics.cu -- label post-processing + MoF counters on the device.
//
// Replaces, for the Viterbi head, reference src/core/utils.py:34-47 (make_same_size_interpolate:
// nearest-neighbour resize of the predicted label vector to the ground-truth length) followed by
// src/core/metrics/segmentation.py:16-44 (MoFAccuracyMetric.add: frames whose target is not in
// ignore_ids; correct = target == prediction), as called at src/mucon/evaluators.py:225-243.
// Only two counters per video cross the bus instead of the label vectors.
#include "common.cuh"

namespace mucon {
namespace {

struct IgnoreIds {
  int n;
  int id[16];
};

__global__ void __launch_bounds__(256) vit_mof_kernel(const int32_t* __restrict__ pred, const int64_t* __restrict__ pred_off,
                                                      const int32_t* __restrict__ gt, const int64_t* __restrict__ gt_off,
                                                      IgnoreIds ign, unsigned long long* __restrict__ counts) {
  const int v = blockIdx.y;
  const int64_t p0 = pred_off[v], g0 = gt_off[v];
  const int Tp = static_cast<int>(pred_off[v + 1] - p0);
  const int Tg = static_cast<int>(gt_off[v + 1] - g0);
  if (Tp <= 0 || Tg <= 0) return;
  // F.interpolate(mode="nearest"): src = min(floor(dst * (float)in / out), in - 1)
  const float scale = static_cast<float>(Tp) / static_cast<float>(Tg);
  unsigned correct = 0, total = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Tg; i += gridDim.x * blockDim.x) {
    int src = static_cast<int>(floorf(static_cast<float>(i) * scale));
    if (src > Tp - 1) src = Tp - 1;
    const int t = gt[g0 + i];
    bool skip = false;
    for (int k = 0; k < ign.n; ++k) skip |= (t == ign.id[k]);
    if (!skip) {
      ++total;
      correct += (pred[p0 + src] == t);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    correct += __shfl_xor_sync(0xffffffffu, correct, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
  }
  if ((threadIdx.x & 31) == 0 && total) {
    atomicAdd(&counts[2 * v], static_cast<unsigned long long>(correct));
    atomicAdd(&counts[2 * v + 1], static_cast<unsigned long long>(total));
  }
}


// ---------------------------------------------------------------------------------------------
// Segment-level metrics of the Viterbi head, one CTA per video: the predicted labels are resized to the ground
// truth's length (make_same_size_interpolate, src/core/utils.py:34-47), both label vectors are cut into maximal
// runs whose label is not in ignore_ids (isba_code.py:11-21,36-44; mstcn_code.py:6-27), then
//   IoD / IoU   isba_code.py:22-61 / 64-109: mean over true runs of the best same-label score
//   Edit        mstcn_code.py:30-56: (1 - levenshtein / max(m, n)) * 100 over the two label sequences
//   F1 counts   mstcn_code.py:59-81: greedy IoU matching at overlaps 0.1 / 0.25 / 0.5 -> tp, fp, fn
// out[v] = {iod, iou, edit, tp, fp, fn (x3)} as doubles.  Ratios are IEEE double divisions of the same integers
// the reference divides, so only the summation order of the two means can differ from NumPy's.
constexpr int kSegThreads = 256;
constexpr int kSegWsPerFrame = 18;  // int32 words of workspace per (ground-truth frame + 1)

struct RunList {
  int* lab;
  int* st;
  int* en;
  int n;
};

__device__ __forceinline__ bool ignored(const IgnoreIds& ign, int l) {
  bool s = false;
  for (int k = 0; k < ign.n; ++k) s |= (l == ign.id[k]);
  return s;
}

// exclusive prefix sum of one int per thread (kSegThreads threads); returns the total
__device__ int block_exclusive_scan(int v, int* sm, int& total) {
  sm[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < kSegThreads; ++i) { const int x = sm[i]; sm[i] = acc; acc += x; }
    sm[kSegThreads] = acc;
  }
  __syncthreads();
  const int r = sm[threadIdx.x];
  total = sm[kSegThreads];
  __syncthreads();
  return r;
}

// label of frame i of the sequence being cut: the ground truth, or the prediction resized to its length
template <bool RESIZED>
__device__ __forceinline__ int frame_label(const int32_t* base, int i, int Tsrc, float scale) {
  if (!RESIZED) return base[i];
  int src = static_cast<int>(floorf(static_cast<float>(i) * scale));
  if (src > Tsrc - 1) src = Tsrc - 1;
  return base[src];
}

template <bool RESIZED>
__device__ void cut_runs(const int32_t* base, int T, int Tsrc, float scale, const IgnoreIds& ign, int* tmpL, int* tmpS,
                         RunList& out, int* sm) {
  // pass 1: all maximal runs (label, start), frames split into one contiguous chunk per thread
  const int per = (T + kSegThreads - 1) / kSegThreads;
  const int a = min(T, static_cast<int>(threadIdx.x) * per), b = min(T, a + per);
  int cnt = 0;
  for (int i = a; i < b; ++i) {
    const int l = frame_label<RESIZED>(base, i, Tsrc, scale);
    if (i == 0 || l != frame_label<RESIZED>(base, i - 1, Tsrc, scale)) ++cnt;
  }
  int nall;
  int pos = block_exclusive_scan(cnt, sm, nall);
  for (int i = a; i < b; ++i) {
    const int l = frame_label<RESIZED>(base, i, Tsrc, scale);
    if (i == 0 || l != frame_label<RESIZED>(base, i - 1, Tsrc, scale)) { tmpL[pos] = l; tmpS[pos] = i; ++pos; }
  }
  __syncthreads();
  // pass 2: keep the runs whose label is not ignored; a run ends where the next run (kept or not) starts
  const int per2 = (nall + kSegThreads - 1) / kSegThreads;
  const int a2 = min(nall, static_cast<int>(threadIdx.x) * per2), b2 = min(nall, a2 + per2);
  cnt = 0;
  for (int k = a2; k < b2; ++k) cnt += ignored(ign, tmpL[k]) ? 0 : 1;
  pos = block_exclusive_scan(cnt, sm, out.n);
  for (int k = a2; k < b2; ++k) {
    if (ignored(ign, tmpL[k])) continue;
    out.lab[pos] = tmpL[k];
    out.st[pos] = tmpS[k];
    out.en[pos] = (k + 1 < nall) ? tmpS[k + 1] : T;
    ++pos;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSegThreads) vit_segment_metrics_kernel(
    const int32_t* __restrict__ pred, const int64_t* __restrict__ pred_off, const int32_t* __restrict__ gt,
    const int64_t* __restrict__ gt_off, IgnoreIds ign, int32_t* __restrict__ ws, double* __restrict__ out) {
  __shared__ int sm[kSegThreads + 1];
  __shared__ double rv[kSegThreads];
  __shared__ int ri[kSegThreads];
  __shared__ double res[12];
  const int v = blockIdx.x;
  const int64_t p0 = pred_off[v], g0 = gt_off[v];
  const int Tp = static_cast<int>(pred_off[v + 1] - p0);
  const int Tg = static_cast<int>(gt_off[v + 1] - g0);
  double* o = out + static_cast<int64_t>(v) * 12;
  if (Tp <= 0 || Tg <= 0) {
    if (threadIdx.x < 12) o[threadIdx.x] = threadIdx.x < 3 ? nan("") : 0.0;
    return;
  }
  const int cap = Tg + 1;
  int32_t* w = ws + kSegWsPerFrame * (g0 + v);
  double* sc_iod = reinterpret_cast<double*>(w);            // [cap]
  double* sc_iou = sc_iod + cap;                             // [cap]
  int* tmpL = w + 4 * cap;
  int* tmpS = w + 5 * cap;
  RunList Y{w + 6 * cap, w + 7 * cap, w + 8 * cap, 0}, P{w + 9 * cap, w + 10 * cap, w + 11 * cap, 0};
  int* diag = w + 12 * cap;   // three rolling anti-diagonals of the edit-distance table, [cap] each
  int* hits = w + 15 * cap;   // [3][cap]
  const float scale = static_cast<float>(Tp) / static_cast<float>(Tg);
  cut_runs<false>(gt + g0, Tg, Tg, 1.f, ign, tmpL, tmpS, Y, sm);
  cut_runs<true>(pred + p0, Tg, Tp, scale, ign, tmpL, tmpS, P, sm);
  const int ny = Y.n, np_ = P.n;

  // ---- IoD / IoU: best same-label score per true run
  for (int i = threadIdx.x; i < ny; i += kSegThreads) {
    double bd = 0.0, bu = 0.0;
    const int yl = Y.lab[i], ys = Y.st[i], ye = Y.en[i];
    for (int j = 0; j < np_; ++j) {
      if (P.lab[j] != yl) continue;
      const int ps = P.st[j], pe = P.en[j];
      const double inter = static_cast<double>(min(pe, ye) - max(ps, ys));
      bd = fmax(bd, inter / static_cast<double>(pe - ps));
      bu = fmax(bu, inter / static_cast<double>(max(pe, ye) - min(ps, ys)));
    }
    sc_iod[i] = bd;
    sc_iou[i] = bu;
  }
  for (int i = threadIdx.x; i < 3 * cap; i += kSegThreads) hits[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < ny; ++i) { a += sc_iod[i]; b += sc_iou[i]; }
    res[0] = ny ? a / ny : nan("");
    res[1] = ny ? b / ny : nan("");
  }

  // ---- edit distance between the two label sequences (rows = predicted runs, columns = true runs), by anti-diagonals
  {
    const int m = np_, n = ny;
    int* d2 = diag;            // diagonal d - 2
    int* d1 = diag + cap;      // diagonal d - 1
    int* d0 = diag + 2 * cap;  // diagonal d
    for (int d = 0; d <= m + n; ++d) {
      const int ilo = max(0, d - n), ihi = min(m, d);
      for (int i = ilo + threadIdx.x; i <= ihi; i += kSegThreads) {
        const int j = d - i;
        int val;
        if (i == 0) val = j;
        else if (j == 0) val = i;
        else if (P.lab[i - 1] == Y.lab[j - 1]) val = d2[i - 1];
        else val = min(min(d1[i - 1], d1[i]), d2[i - 1]) + 1;
        d0[i] = val;
      }
      __syncthreads();
      int* t = d2; d2 = d1; d1 = d0; d0 = t;
    }
    // after the last rotation d1 holds diagonal m + n
    if (threadIdx.x == 0) {
      const int mx = max(m, n);
      res[2] = mx ? (1.0 - static_cast<double>(d1[m]) / mx) * 100.0 : nan("");
    }
  }

  // ---- F1 counts: predicted runs in order, each matched to the true run of highest IoU (first on ties)
  double tp[3] = {0, 0, 0}, fp[3] = {0, 0, 0};
  const double ov[3] = {0.1, 0.25, 0.5};
  for (int j = 0; j < np_; ++j) {
    double best = -INFINITY;
    int bi = 0x7fffffff;
    const int pl = P.lab[j], ps = P.st[j], pe = P.en[j];
    for (int i = threadIdx.x; i < ny; i += kSegThreads) {
      const int ys = Y.st[i], ye = Y.en[i];
      double val = static_cast<double>(min(pe, ye) - max(ps, ys)) / static_cast<double>(max(pe, ye) - min(ps, ys));
      val = val * (pl == Y.lab[i] ? 1.0 : 0.0);
      if (val > best) { best = val; bi = i; }  // i ascends within a thread: the first maximum stays
    }
    rv[threadIdx.x] = best;
    ri[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int t = 1; t < kSegThreads; ++t)
        if (rv[t] > best || (rv[t] == best && ri[t] < bi)) { best = rv[t]; bi = ri[t]; }
      for (int k = 0; k < 3; ++k) {
        if (ny > 0 && best >= ov[k] && !hits[k * cap + bi]) { tp[k] += 1.0; hits[k * cap + bi] = 1; }
        else fp[k] += 1.0;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) {
      int h = 0;
      for (int i = 0; i < ny; ++i) h += hits[k * cap + i];
      res[3 + 3 * k] = tp[k];
      res[4 + 3 * k] = fp[k];
      res[5 + 3 * k] = static_cast<double>(ny - h);
    }
  }
  __syncthreads();
  if (threadIdx.x < 12) o[threadIdx.x] = res[threadIdx.x];
}

}  // namespace
}  // namespace mucon

using namespace mucon;

extern "C" int mucon_vit_mof(const int32_t* pred, const int64_t* pred_off, const int32_t* gt, const int64_t* gt_off,
                             int V, int max_T_gt, const int32_t* ignore_ids_h, int n_ignore,
                             unsigned long long* counts, void* stream) {
  if (!pred || !pred_off || !gt || !gt_off || !counts || V < 0 || max_T_gt < 0 || n_ignore < 0) return MUCON_EINVAL;
  if (n_ignore > 16 || (n_ignore && !ignore_ids_h)) return MUCON_EUNSUPPORTED;
  if (V == 0) return MUCON_OK;
  if (V > 65535) return MUCON_EUNSUPPORTED;
  IgnoreIds ign;
  ign.n = n_ignore;
  for (int k = 0; k < 16; ++k) ign.id[k] = k < n_ignore ? ignore_ids_h[k] : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MUCON_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 2 * V, st));
  if (max_T_gt == 0) return MUCON_OK;
  int bx = (max_T_gt + 255) / 256;
  if (bx > 16) bx = 16;
  vit_mof_kernel<<<dim3(bx, V), 256, 0, st>>>(pred, pred_off, gt, gt_off, ign, counts);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_vit_segment_metrics(const int32_t* pred, const int64_t* pred_off, const int32_t* gt,
                                         const int64_t* gt_off, int V, const int32_t* ignore_ids_h, int n_ignore,
                                         int32_t* ws, double* out, void* stream) {
  if (!pred || !pred_off || !gt || !gt_off || !ws || !out || V < 0 || n_ignore < 0) return MUCON_EINVAL;
  if (n_ignore > 16 || (n_ignore && !ignore_ids_h)) return MUCON_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(ws) & 7) != 0) return MUCON_EALIGN;
  if (V == 0) return MUCON_OK;
  IgnoreIds ign;
  ign.n = n_ignore;
  for (int k = 0; k < 16; ++k) ign.id[k] = k < n_ignore ? ignore_ids_h[k] : 0;
  vit_segment_metrics_kernel<<<V, kSegThreads, 0, static_cast<cudaStream_t>(stream)>>>(pred, pred_off, gt, gt_off, ign, ws,
                                                                                      out);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int64_t mucon_vit_segment_metrics_ws_words(int64_t total_gt_frames, int V) {
  return kSegWsPerFrame * (total_gt_frames + V) + 2;
}
