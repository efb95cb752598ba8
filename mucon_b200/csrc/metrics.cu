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
