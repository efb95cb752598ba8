// backbone.cu -- forward pass of the dilated temporal-conv backbone (sm_100a).
//
// Replaces, for inference, reference src/core/modules/temporal.py:43-53,128-147 (WaveNetLayer /
// WaveNetBlock.forward), src/mucon/models.py:759-768 (GroupNorm + ReLU tail), :567-582 (nearest
// upsample + 1x1 classifier) and :368 (log_softmax).  Activations are time-major, channels
// contiguous: a video is a [T, C] block of rows, videos are concatenated (row offsets per
// resolution), so a 1x1 conv is a GEMM over rows and the classifier can run at the pooled
// resolution before the nearest-neighbour expansion (a 1x1 conv commutes with it).
//
//   mucon_gemm_tf32_bias_act      2048 -> 128 input projection on tcgen05 (backbone_gemm.cuh)
//   mucon_conv1d                  k = 1 / k = 3 dilated conv, fused bias / ReLU / residual (fp32 FFMA)
//   mucon_maxpool2                max_pool1d(2), floor
//   mucon_groupnorm_relu          GroupNorm over (channels of the group x time) per video, + ReLU
//   mucon_logsoftmax_expand       log_softmax over classes at pooled resolution, expanded to T frames
#include <cuda.h>
#include <math.h>

#include "backbone_gemm.cuh"
#include "backbone_bf16.cuh"
#include "backbone_train.cuh"
#include <stdlib.h>

#include "common.cuh"

namespace mucon {
namespace {

// ---------------------------------------------------------------------------------------------
// tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, cols] fp32 row-major, box [box_rows, 32 cols] with 128-byte swizzle
int make_map_2d(CUtensorMap* m, const float* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return MUCON_ECUDA;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MUCON_OK : MUCON_ECUDA;
}

// [rows, 128] bf16 row-major, box [box_rows, 64 cols] (128 bytes) with 128-byte swizzle
int make_map_bf16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, bool f16 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return MUCON_ECUDA;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MUCON_OK : MUCON_ECUDA;
}

// ---------------------------------------------------------------------------------------------
// k = 1 / k = 3 dilated conv on CUDA cores: out[t, co] = b[co] + sum_tap sum_ci W[tap][ci][co] * x[t + (tap-c)*dil, ci]
// with zero padding at the video's ends (Conv1d(padding = dilation), temporal.py:21-27).
// CTA: 64 rows x 64 output channels, 256 threads, 4 x 4 outputs per thread.
template <int TAPS>
__global__ void __launch_bounds__(256) conv1d_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                     const float* __restrict__ W, const float* __restrict__ bias,
                                                     const float* __restrict__ residual,
                                                     const int64_t* __restrict__ off, int Cin, int Cout, int dil,
                                                     int relu_in, int relu_out) {
  __shared__ float Xs[64][33];
  __shared__ float Ws[32][64];
  const int v = blockIdx.y;
  const int64_t r0 = off[v];
  const int T = static_cast<int>(off[v + 1] - r0);
  const int t0 = blockIdx.x * 64;
  if (t0 >= T) return;
  const int co0 = blockIdx.z * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < TAPS; ++tap) {
    const int shift = (tap - TAPS / 2) * dil;
    // a tap that only ever sees padding contributes nothing (dilation >= T: temporal.py layers 8-10)
    if (shift >= T || -shift >= T) continue;
    for (int ci0 = 0; ci0 < Cin; ci0 += 32) {
      for (int i = tid; i < 64 * 32; i += 256) {
        const int r = i >> 5, c = i & 31;
        const int t = t0 + r + shift;
        float x = 0.f;
        if (t >= 0 && t < T && ci0 + c < Cin) {
          x = in[(r0 + t) * Cin + ci0 + c];
          x = act_mode(x, relu_in);
        }
        Xs[r][c] = x;
      }
      for (int i = tid; i < 32 * 64; i += 256) {
        const int c = i >> 6, o = i & 63;
        Ws[c][o] = (ci0 + c < Cin && co0 + o < Cout) ? W[(static_cast<size_t>(tap) * Cin + ci0 + c) * Cout + co0 + o] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int c = 0; c < 32; ++c) {
        float xv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = Xs[ty * 4 + i][c];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = Ws[c][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= T) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= Cout) continue;
      float y = acc[i][j] + bias[co];
      y = act_mode(y, relu_out);
      if (residual) y += residual[(r0 + t) * Cout + co];
      out[(r0 + t) * Cout + co] = y;
    }
  }
}

// mode 0: max_pool1d(2); mode 1: avg_pool1d(2) * 2 = sum of the pair (temporal.py:139-142)
__global__ void maxpool2_kernel(const float* __restrict__ in, float* __restrict__ out, const int64_t* __restrict__ off_in,
                                const int64_t* __restrict__ off_out, int C, int mode) {
  const int v = blockIdx.y;
  const int64_t i0 = off_in[v], o0 = off_out[v];
  const int To = static_cast<int>(off_out[v + 1] - o0);
  const int64_t n = static_cast<int64_t>(To) * C;
  if ((C & 3) == 0) {  // 16-byte path: rows are multiples of 16 bytes
    const int c4 = C >> 2;
    const int n4 = static_cast<int>(n >> 2);
    const float4* in4 = reinterpret_cast<const float4*>(in + i0 * C);
    float4* out4 = reinterpret_cast<float4*>(out + o0 * C);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      const int t = i / c4, q = i - t * c4;
      const float4 a = in4[(2 * t) * c4 + q], b = in4[(2 * t + 1) * c4 + q];
      out4[i] = mode ? make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w)
                     : make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
    }
    return;
  }
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / C;
    const int c = static_cast<int>(i - t * C);
    const float a = in[(i0 + 2 * t) * C + c], b = in[(i0 + 2 * t + 1) * C + c];
    out[o0 * C + i] = mode ? a + b : fmaxf(a, b);
  }
}

// GroupNorm(groups, C) over one video's [T, C] block (statistics over T x C/groups), optional ReLU.
__global__ void __launch_bounds__(256) groupnorm_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const int64_t* __restrict__ off, int C, int groups, float eps,
                                                        int relu) {
  extern __shared__ double red[];  // [2][C]
  const int v = blockIdx.x;
  const int64_t r0 = off[v];
  const int T = static_cast<int>(off[v + 1] - r0);
  double* s1 = red;
  double* s2 = red + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int t = 0; t < T; ++t) {
      const double x = in[(r0 + t) * C + c];
      a += x;
      b += x * x;
    }
    s1[c] = a;
    s2[c] = b;
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    double a = 0.0, b = 0.0;
    for (int k = 0; k < cpg; ++k) { a += s1[g * cpg + k]; b += s2[g * cpg + k]; }
    const double n = static_cast<double>(T) * cpg;
    const double mean = a / n;
    double var = b / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float mu = static_cast<float>(mean);
    const float ga = gamma[c], be = beta[c];
    for (int t = 0; t < T; ++t) {
      float y = (in[(r0 + t) * C + c] - mu) * rstd * ga + be;
      if (relu) y = fmaxf(y, 0.f);
      out[(r0 + t) * C + c] = y;
    }
  }
}

// log_softmax over C classes of the pooled-resolution logits row idx(t), written for every frame t:
// idx(t) = min(floor(t * (float)Tz / T), Tz - 1) -- F.interpolate(mode="nearest") (models.py:574-576).
// A CTA owns kLseFrames consecutive frames: its warps first turn the few source rows those frames
// map to into log-probabilities in shared memory, then all threads stream them out with coalesced
// 16-byte stores (the output, 4*C bytes per frame, is the only large traffic of this kernel).
constexpr int kLseFrames = 256;
constexpr int kLseMaxRows = 64;
__global__ void __launch_bounds__(256) logsoftmax_expand_kernel(const float* __restrict__ logits,
                                                                const int64_t* __restrict__ off_z,
                                                                const int64_t* __restrict__ off_t, int C,
                                                                float* __restrict__ out) {
  extern __shared__ float lrows[];  // [kLseMaxRows][C]
  const int v = blockIdx.y;
  const int64_t z0 = off_z[v], t0v = off_t[v];
  const int Tz = static_cast<int>(off_z[v + 1] - z0);
  const int T = static_cast<int>(off_t[v + 1] - t0v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float scale = static_cast<float>(Tz) / static_cast<float>(T);
  auto src_row = [&](int t) {
    int iz = static_cast<int>(floorf(static_cast<float>(t) * scale));
    return iz > Tz - 1 ? Tz - 1 : iz;
  };
  for (int t0 = blockIdx.x * kLseFrames; t0 < T; t0 += gridDim.x * kLseFrames) {
    const int t1 = min(T, t0 + kLseFrames);
    const int iz0 = src_row(t0), iz1 = src_row(t1 - 1);
    const int nrows = iz1 - iz0 + 1;
    if (nrows <= kLseMaxRows) {
      for (int r = warp; r < nrows; r += nwarp) {
        const float* row = logits + (z0 + iz0 + r) * C;
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float lse = m + logf(s);
        for (int c = lane; c < C; c += 32) lrows[r * C + c] = row[c] - lse;
      }
      __syncthreads();
      float* obase = out + (t0v + t0) * C;
      if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0) {
        const int c4 = C >> 2;
        const int n = (t1 - t0) * c4;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const int t = i / c4, q = i - t * c4;
          const int r = src_row(t0 + t) - iz0;
          reinterpret_cast<float4*>(obase)[i] = *reinterpret_cast<const float4*>(lrows + r * C + 4 * q);
        }
      } else {
        const int n = (t1 - t0) * C;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const int t = i / C, c = i - t * C;
          obase[i] = lrows[(src_row(t0 + t) - iz0) * C + c];
        }
      }
      __syncthreads();
    } else {
      // down-sampling or nearly 1:1 mapping: one warp per frame
      for (int t = t0 + warp; t < t1; t += nwarp) {
        const float* row = logits + (z0 + src_row(t)) * C;
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float lse = m + logf(s);
        float* orow = out + (t0v + t) * C;
        for (int c = lane; c < C; c += 32) orow[c] = row[c] - lse;
      }
    }
  }
}

}  // namespace
}  // namespace mucon

using namespace mucon;

static int launch_proj(const float* A, int64_t M, int K, const float* W, int N, const float* bias, void* out, int relu,
                       int out_bf16, void* stream);

extern "C" int mucon_gemm_tf32_bias_act(const float* A, int64_t M, int K, const float* W, int N, const float* bias,
                                        float* out, int relu, void* stream) {
  return launch_proj(A, M, K, W, N, bias, out, relu, 0, stream);
}

extern "C" int mucon_gemm_tf32_bias_act_bf16(const float* A, int64_t M, int K, const float* W, int N,
                                             const float* bias, void* out16, int relu, int fp16, void* stream) {
  return launch_proj(A, M, K, W, N, bias, out16, relu, fp16 ? 2 : 1, stream);
}

static int launch_proj(const float* A, int64_t M, int K, const float* W, int N, const float* bias, void* out, int relu,
                       int out_bf16, void* stream) {
  if (!A || !W || !bias || !out || M < 0 || K < 1) return MUCON_EINVAL;
  if (N != gemm::BN || K % gemm::BK != 0) return MUCON_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 31))
    return MUCON_EALIGN;
  if (M == 0) return MUCON_OK;
  if (M > 0x7fffffff) return MUCON_EUNSUPPORTED;
  const int sms = mucon_device_sm_count();   // of the current device (cached per device)
  // 256-row CTA tiles (one weight stage feeds two M tiles) once there are enough of them to fill the GPU
  static const int force_mt = getenv("MUCON_PROJ_MT") ? atoi(getenv("MUCON_PROJ_MT")) : 0;
  const int mt = force_mt ? force_mt : (M >= static_cast<int64_t>(2) * gemm::BM * sms ? 2 : 1);
  CUtensorMap ta, tb;
  int rc = make_map_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), mt * gemm::BM);
  if (rc != MUCON_OK) return rc;
  rc = make_map_2d(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), gemm::BN);
  if (rc != MUCON_OK) return rc;
  const int tile_m = mt * gemm::BM;
  const int tiles = static_cast<int>((M + tile_m - 1) / tile_m);
  const int grid = tiles < sms ? tiles : sms;
  if (mt == 2) {
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(gemm::proj_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          gemm::SMEM_BYTES));
    gemm::proj_gemm_kernel<2><<<grid, gemm::THREADS, gemm::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(
        ta, tb, bias, static_cast<float*>(out), static_cast<int>(M), K, relu, out_bf16);
  } else {
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(gemm::proj_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          gemm::SMEM_BYTES));
    gemm::proj_gemm_kernel<1><<<grid, gemm::THREADS, gemm::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(
        ta, tb, bias, static_cast<float*>(out), static_cast<int>(M), K, relu, out_bf16);
  }
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

static int launch_conv_gemm(const float* in, float* out, const float* W_kco, const float* bias, const float* residual,
                            const void* tiles, int num_tiles, int64_t rows, const convgemm::TapShifts& ts,
                            int relu_mid, int relu_final, void* stream, const float* mul = nullptr,
                            const float* gate = nullptr, bool bias_optional = false) {
  if (!in || !out || !W_kco || (!bias && !bias_optional) || !tiles || num_tiles < 0 || rows < 0) return MUCON_EINVAL;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(W_kco) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (residual && (reinterpret_cast<uintptr_t>(residual) & 15)) ||
      (mul && (reinterpret_cast<uintptr_t>(mul) & 15)) || (gate && (reinterpret_cast<uintptr_t>(gate) & 15)))
    return MUCON_EALIGN;
  if (num_tiles == 0 || rows == 0) return MUCON_OK;
  if (rows > 0x7fffffff - 4096) return MUCON_EUNSUPPORTED;
  CUtensorMap tx, tw;
  int rc = make_map_2d(&tx, in, static_cast<uint64_t>(rows), convgemm::C, gemm::BM);
  if (rc != MUCON_OK) return rc;
  rc = make_map_2d(&tw, W_kco, static_cast<uint64_t>(ts.n) * convgemm::C, convgemm::C, gemm::BN);
  if (rc != MUCON_OK) return rc;
  const int sms = mucon_device_sm_count();   // of the current device (cached per device)
  const int grid = num_tiles < sms ? num_tiles : sms;
  MUCON_CUDA_CHECK(cudaFuncSetAttribute(convgemm::conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        convgemm::CSMEM_BYTES));
  convgemm::conv_gemm_kernel<<<grid, convgemm::CTHREADS, convgemm::CSMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(
      tx, tw, static_cast<const convgemm::Tile*>(tiles), num_tiles, ts, bias, residual, out, relu_mid, relu_final, mul,
      gate);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_conv_gemm_tf32(const float* in, float* out, const float* W_kco, const float* bias,
                                    const float* residual, const void* tiles, int num_tiles, int64_t rows, int taps,
                                    int dilation, int relu_mid, int relu_final, void* stream) {
  if (dilation < 1) return MUCON_EINVAL;
  if (taps != 1 && taps != 3) return MUCON_EUNSUPPORTED;
  convgemm::TapShifts ts{};
  ts.n = taps;
  for (int t = 0; t < taps; ++t) ts.s[t] = (t - taps / 2) * dilation;
  return launch_conv_gemm(in, out, W_kco, bias, residual, tiles, num_tiles, rows, ts, relu_mid, relu_final, stream);
}

extern "C" int mucon_conv_gemm_tf32_shifts(const float* in, float* out, const float* W_kco, const float* bias,
                                           const float* residual, const void* tiles, int num_tiles, int64_t rows,
                                           const int32_t* shifts_h, int n_shifts, int relu_mid, int relu_final,
                                           void* stream) {
  if (!shifts_h || n_shifts < 1) return MUCON_EINVAL;
  if (n_shifts > convgemm::kMaxTaps) return MUCON_EUNSUPPORTED;
  convgemm::TapShifts ts{};
  ts.n = n_shifts;
  bool has_zero = false;
  for (int t = 0; t < n_shifts; ++t) {
    ts.s[t] = shifts_h[t];
    has_zero |= shifts_h[t] == 0;
  }
  if (!has_zero) return MUCON_EINVAL;  // the kernel relies on one tap that is live for every tile
  return launch_conv_gemm(in, out, W_kco, bias, residual, tiles, num_tiles, rows, ts, relu_mid, relu_final, stream);
}

extern "C" int mucon_conv_gemm_tf32_ex(const float* in, float* out, const float* W_kco, const float* bias,
                                       const float* residual, const float* mul, const float* gate, const void* tiles,
                                       int num_tiles, int64_t rows, const int32_t* shifts_h, int n_shifts, int relu_mid,
                                       int relu_final, void* stream) {
  if (!shifts_h || n_shifts < 1) return MUCON_EINVAL;
  if (n_shifts > convgemm::kMaxTaps) return MUCON_EUNSUPPORTED;
  convgemm::TapShifts ts{};
  ts.n = n_shifts;
  bool has_zero = false;
  for (int t = 0; t < n_shifts; ++t) {
    ts.s[t] = shifts_h[t];
    has_zero |= shifts_h[t] == 0;
  }
  if (!has_zero) return MUCON_EINVAL;
  return launch_conv_gemm(in, out, W_kco, bias, residual, tiles, num_tiles, rows, ts, relu_mid, relu_final, stream, mul,
                          gate, true);
}

// ---------------------------------------------------------------------------------------------
// training step: weight / bias gradients (backbone_train.cuh) and the max-pool backward
extern "C" int mucon_wgrad_tf32(const float* dY, const float* X, int ldx, const void* tiles, int num_tiles, int64_t rows,
                                const int32_t* shifts_h, const int32_t* xcol_h, const int64_t* out_off_h, int n_jobs,
                                int ldo, float* dW, float* dbias, void* stream) {
  if (!dY || !X || !tiles || !shifts_h || !xcol_h || !out_off_h || !dW || num_tiles < 0 || rows < 0 || n_jobs < 1 ||
      ldx < 128 || ldo < 128)
    return MUCON_EINVAL;
  if (n_jobs > wgrad::kMaxJobs || ldx % 32 != 0 || ldo % 4 != 0) return MUCON_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(dY) & 15) || (reinterpret_cast<uintptr_t>(X) & 15) ||
      (reinterpret_cast<uintptr_t>(dW) & 15))
    return MUCON_EALIGN;
  if (num_tiles == 0 || rows == 0) return MUCON_OK;
  if (rows > 0x7fffffff - 4096) return MUCON_EUNSUPPORTED;
  wgrad::Jobs jobs{};
  jobs.n = n_jobs;
  jobs.ldo = ldo;
  bool zero_in_group0 = false;
  for (int j = 0; j < n_jobs; ++j) {
    if (xcol_h[j] < 0 || xcol_h[j] % 32 != 0 || xcol_h[j] + 128 > ldx || (out_off_h[j] & 3)) return MUCON_EINVAL;
    jobs.j[j].shift = shifts_h[j];
    jobs.j[j].xblk = xcol_h[j] / 32;
    jobs.j[j].out_off = out_off_h[j];
    if (j < wgrad::kJobsPerCta && shifts_h[j] == 0) zero_in_group0 = true;
  }
  if (dbias && !zero_in_group0) return MUCON_EINVAL;  // the bias sum rides on stages that are always live
  CUtensorMap tdy, tx;
  int rc = make_map_2d(&tdy, dY, static_cast<uint64_t>(rows), wgrad::C, wgrad::KT, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != MUCON_OK) return rc;
  rc = make_map_2d(&tx, X, static_cast<uint64_t>(rows), static_cast<uint64_t>(ldx), wgrad::KT,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != MUCON_OK) return rc;
  const int groups = (n_jobs + wgrad::kJobsPerCta - 1) / wgrad::kJobsPerCta;
  const int nj_max = n_jobs < wgrad::kJobsPerCta ? n_jobs : wgrad::kJobsPerCta;
  const int stage_bytes = (1 + nj_max) * wgrad::OP_BYTES;
  int stages = (wgrad::WSMEM_MAX - 2048) / stage_bytes;
  if (stages > wgrad::kMaxStages) stages = wgrad::kMaxStages;
  const int smem = 1024 + stages * stage_bytes + 512;
  const int sms = mucon_device_sm_count();
  int gx = sms / groups;
  if (gx < 1) gx = 1;
  if (gx > num_tiles) gx = num_tiles;
  MUCON_CUDA_CHECK(cudaFuncSetAttribute(wgrad::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  wgrad::wgrad_kernel<<<dim3(gx, groups), wgrad::WTHREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      tdy, tx, static_cast<const convgemm::Tile*>(tiles), num_tiles, jobs, stages, dW, dbias);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_maxpool2_bwd(const float* x, const float* dy, const int64_t* off_in, const int64_t* off_out, int V,
                                  int max_T_out, int C, float* dx, void* stream) {
  if (!x || !dy || !off_in || !off_out || !dx || V < 0 || C < 4 || max_T_out < 0) return MUCON_EINVAL;
  if (C % 4 != 0 || V > 65535) return MUCON_EUNSUPPORTED;
  if (V == 0) return MUCON_OK;
  int bx = static_cast<int>((static_cast<int64_t>(max_T_out) * (C / 4) + 255) / 256);
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  wgrad::maxpool2_bwd_kernel<<<dim3(bx, V), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, off_in, off_out, C, dx);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

static int launch_layer(const float* x, float* out, const float* Wd_kco, const float* bd, const float* W1_kco,
                        const float* b1, const void* tiles, int num_tiles, int64_t rows, int dilation, int pool,
                        int relu_final, bool pair, void* stream) {
  if (!x || !out || !Wd_kco || !bd || !W1_kco || !b1 || !tiles || num_tiles < 0 || rows < 0 || dilation < 1)
    return MUCON_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(Wd_kco) & 15) ||
      (reinterpret_cast<uintptr_t>(W1_kco) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return MUCON_EALIGN;
  if (pair && (num_tiles & 1)) return MUCON_EINVAL;
  if (num_tiles == 0 || rows == 0) return MUCON_OK;
  if (rows > 0x7fffffff - 4096) return MUCON_EUNSUPPORTED;
  CUtensorMap tx, twd, tw1;
  const uint32_t wbox = pair ? gemm::BN / 2 : gemm::BN;  // paired CTAs load half of every weight k-block each
  int rc = make_map_2d(&tx, x, static_cast<uint64_t>(rows), layer::C, gemm::BM);
  if (rc != MUCON_OK) return rc;
  rc = make_map_2d(&twd, Wd_kco, 3ull * layer::C, layer::C, wbox);
  if (rc != MUCON_OK) return rc;
  rc = make_map_2d(&tw1, W1_kco, layer::C, layer::C, wbox);
  if (rc != MUCON_OK) return rc;
  const int sms = mucon_device_sm_count();   // of the current device (cached per device)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const layer::Tile* tl = static_cast<const layer::Tile*>(tiles);
  static int use_slab = -1;
  if (use_slab < 0) {
    const char* e = getenv("MUCON_LAYER_SLAB");
    use_slab = (e && e[0] == '0') ? 0 : 1;
  }
  if (!pair && use_slab && dilation <= layer::kSlabMaxDil) {
    // small dilations: one activation slab of 128 + 2*dilation rows per k-block serves all three taps
    CUtensorMap txs;
    rc = make_map_2d(&txs, x, static_cast<uint64_t>(rows), layer::C, gemm::BM + 2 * dilation);
    if (rc != MUCON_OK) return rc;
    const int grid = num_tiles < sms ? num_tiles : sms;
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(layer::wavenet_layer_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          layer::SSMEM_BYTES));
    layer::wavenet_layer_slab_kernel<<<grid, layer::LTHREADS, layer::SSMEM_BYTES, st>>>(
        txs, twd, tw1, tl, num_tiles, dilation, bd, b1, x, out, pool, relu_final);
  } else if (!pair) {
    const int grid = num_tiles < sms ? num_tiles : sms;
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(layer::wavenet_layer_kernel<false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, layer::LSMEM_BYTES));
    layer::wavenet_layer_kernel<false><<<grid, layer::LTHREADS, layer::LSMEM_BYTES, st>>>(
        tx, twd, tw1, tl, num_tiles, dilation, bd, b1, x, out, pool, relu_final);
  } else {
    const int pairs = num_tiles / 2, max_clusters = (sms > 1 ? sms : 2) / 2;
    const int clusters = pairs < max_clusters ? pairs : max_clusters;
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(layer::wavenet_layer_kernel<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, layer::LSMEM_BYTES));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(2 * clusters);
    lc.blockDim = dim3(layer::LTHREADS);
    lc.dynamicSmemBytes = layer::LSMEM_BYTES;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    MUCON_CUDA_CHECK(cudaLaunchKernelEx(&lc, layer::wavenet_layer_kernel<true>, tx, twd, tw1, tl, num_tiles, dilation, bd,
                                        b1, x, out, pool, relu_final));
  }
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_wavenet_layer_tf32(const float* x, float* out, const float* Wd_kco, const float* bd,
                                        const float* W1_kco, const float* b1, const void* tiles, int num_tiles,
                                        int64_t rows, int dilation, int pool, int relu_final, void* stream) {
  return launch_layer(x, out, Wd_kco, bd, W1_kco, b1, tiles, num_tiles, rows, dilation, pool, relu_final, false, stream);
}

extern "C" int mucon_wavenet_layer_tf32_pair(const float* x, float* out, const float* Wd_kco, const float* bd,
                                             const float* W1_kco, const float* b1, const void* tiles, int num_tiles,
                                             int64_t rows, int dilation, int pool, int relu_final, void* stream) {
  return launch_layer(x, out, Wd_kco, bd, W1_kco, b1, tiles, num_tiles, rows, dilation, pool, relu_final, true, stream);
}

extern "C" int mucon_wavenet_layer_bf16_ex(const void* x, void* out, const void* Wd_kco, const float* bd_h,
                                           const void* W1_kco, const float* b1_h, const void* tiles, int num_tiles,
                                           int64_t rows, int64_t rows_out, int dilation, int pool, int relu_final,
                                           int out_f32, int fp16, int residual, void* stream);
extern "C" int mucon_wavenet_layer_bf16(const void* x, void* out, const void* Wd_kco, const float* bd_h,
                                        const void* W1_kco, const float* b1_h, const void* tiles, int num_tiles,
                                        int64_t rows, int64_t rows_out, int dilation, int pool, int relu_final,
                                        int out_f32, int fp16, void* stream) {
  return mucon_wavenet_layer_bf16_ex(x, out, Wd_kco, bd_h, W1_kco, b1_h, tiles, num_tiles, rows, rows_out, dilation, pool,
                                     relu_final, out_f32, fp16, 1, stream);
}

extern "C" int mucon_wavenet_layer_bf16_ex(const void* x, void* out, const void* Wd_kco, const float* bd_h,
                                           const void* W1_kco, const float* b1_h, const void* tiles, int num_tiles,
                                           int64_t rows, int64_t rows_out, int dilation, int pool, int relu_final,
                                           int out_f32, int fp16, int residual, void* stream) {
  if (num_tiles == 0 || rows == 0 || rows_out == 0) return MUCON_OK;  // nothing to produce
  if (!x || !out || !Wd_kco || !bd_h || !W1_kco || !b1_h || !tiles || num_tiles < 0 || rows < 0 || rows_out < 0 ||
      dilation < 1)
    return MUCON_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) & 31) || (reinterpret_cast<uintptr_t>(Wd_kco) & 15) ||
      (reinterpret_cast<uintptr_t>(W1_kco) & 15) || (reinterpret_cast<uintptr_t>(out) & 31))
    return MUCON_EALIGN;
  if (pool && out_f32) return MUCON_EUNSUPPORTED;
  if (rows > 0x7fffffff - 4096) return MUCON_EUNSUPPORTED;
  const int slab = dilation <= layer16::kMaxSlabDil ? 1 : 0;
  CUtensorMap tx, twd, tw1, to;
  const bool f16 = fp16 != 0;
  int rc = make_map_bf16(&tx, x, static_cast<uint64_t>(rows), layer16::C,
                         slab ? gemm::BM + 2 * dilation : gemm::BM, f16);
  if (rc != MUCON_OK) return rc;
  rc = make_map_bf16(&twd, Wd_kco, 3ull * layer16::C, layer16::C, gemm::BN, f16);
  if (rc != MUCON_OK) return rc;
  rc = make_map_bf16(&tw1, W1_kco, layer16::C, layer16::C, gemm::BN, f16);
  if (rc != MUCON_OK) return rc;
  // output tiles leave through a TMA store of the rows that fit the shared-memory staging area
  const int S = layer16::staged_rows(dilation, slab, pool, out_f32);
  if (S > 0) {
    rc = make_map_bf16(&to, out, static_cast<uint64_t>(rows_out), layer16::C, static_cast<uint32_t>(S), f16);
    if (rc != MUCON_OK) return rc;
  } else {
    to = tx;  // never used by the kernel
  }
  layer16::BiasPack bp;
  for (int c = 0; c < layer16::C; ++c) { bp.bd[c] = bd_h[c]; bp.b1[c] = b1_h[c]; }
  const int sms = mucon_device_sm_count();
  const int grid = num_tiles < sms ? num_tiles : sms;
  const int smem = layer16::smem_bytes_of(dilation, slab, pool, out_f32);
  auto kern = f16 ? layer16::wavenet_layer_bf16_kernel<true> : layer16::wavenet_layer_bf16_kernel<false>;
  MUCON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, layer16::SMEM_LIMIT));
  kern<<<grid, layer16::LTHREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      tx, twd, tw1, to, bp, static_cast<const layer16::Tile*>(tiles), num_tiles, dilation, slab, out, pool, relu_final,
      (out_f32 ? 1 : 0) | (residual ? 0 : 2));
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_conv1d(const float* in, float* out, const float* W_tco, const float* bias, const float* residual,
                            const int64_t* row_off, int V, int max_T, int Cin, int Cout, int taps, int dilation,
                            int relu_in, int relu_out, void* stream) {
  if (!in || !out || !W_tco || !bias || !row_off || V < 0 || max_T < 0 || Cin < 1 || Cout < 1 || dilation < 1)
    return MUCON_EINVAL;
  if (taps != 1 && taps != 3) return MUCON_EUNSUPPORTED;
  if (V == 0 || max_T == 0) return MUCON_OK;
  if (V > 65535) return MUCON_EUNSUPPORTED;
  dim3 grid((max_T + 63) / 64, V, (Cout + 63) / 64);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (taps == 1)
    conv1d_kernel<1><<<grid, 256, 0, st>>>(in, out, W_tco, bias, residual, row_off, Cin, Cout, dilation, relu_in, relu_out);
  else
    conv1d_kernel<3><<<grid, 256, 0, st>>>(in, out, W_tco, bias, residual, row_off, Cin, Cout, dilation, relu_in, relu_out);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_pool2(const float* in, float* out, const int64_t* off_in, const int64_t* off_out, int V,
                           int max_T_out, int C, int mode, void* stream);
extern "C" int mucon_maxpool2(const float* in, float* out, const int64_t* off_in, const int64_t* off_out, int V,
                              int max_T_out, int C, void* stream) {
  return mucon_pool2(in, out, off_in, off_out, V, max_T_out, C, 0, stream);
}

extern "C" int mucon_pool2(const float* in, float* out, const int64_t* off_in, const int64_t* off_out, int V,
                           int max_T_out, int C, int mode, void* stream) {
  if (!in || !out || !off_in || !off_out || V < 0 || C < 1 || max_T_out < 0 || mode < 0 || mode > 1) return MUCON_EINVAL;
  if (V == 0 || max_T_out == 0) return MUCON_OK;
  if (V > 65535) return MUCON_EUNSUPPORTED;
  int bx = static_cast<int>((static_cast<int64_t>(max_T_out) * C + 255) / 256);
  if (bx > 64) bx = 64;
  maxpool2_kernel<<<dim3(bx, V), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, off_in, off_out, C, mode);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_groupnorm_relu(const float* in, float* out, const float* gamma, const float* beta,
                                    const int64_t* row_off, int V, int C, int groups, float eps, int relu,
                                    void* stream) {
  if (!in || !out || !gamma || !beta || !row_off || V < 0 || C < 1 || groups < 1 || C % groups) return MUCON_EINVAL;
  if (V == 0) return MUCON_OK;
  groupnorm_kernel<<<V, 128, 2 * C * sizeof(double), static_cast<cudaStream_t>(stream)>>>(in, out, gamma, beta, row_off,
                                                                                         C, groups, eps, relu);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused tail at the pooled resolution (reference src/mucon/models.py:759-768 GroupNorm + ReLU, :574-580 1x1
// classifier -- applied before the nearest upsample, with which it commutes -- and :368 log_softmax):
//   gn_stats_kernel      per video and group: mean and 1/sqrt(var + eps) over (T x C/groups) elements
//   tail_cls_lsm_kernel  per 128-row tile: z = relu(gn(x)), logits = z . Wc^T + bc (fp32 FFMA, k ascending, bias
//                        last: the order of conv1d_kernel), log-probabilities = logits - logsumexp
// Replaces groupnorm_kernel + conv1d_kernel<1> + logsoftmax_rows_kernel (146 + 190 + 20 us on the c2 split).
namespace mucon {
namespace {

constexpr int kTailH = 128;  // hidden channels

__global__ void __launch_bounds__(512) gn_stats_kernel(const float* __restrict__ in, const int64_t* __restrict__ off,
                                                       int groups, float eps, float* __restrict__ stats) {
  __shared__ double s1[4][kTailH], s2[4][kTailH];
  const int v = blockIdx.x;
  const int64_t r0 = off[v];
  const int T = static_cast<int>(off[v + 1] - r0);
  const int c = threadIdx.x & (kTailH - 1), sl = threadIdx.x >> 7;
  double a = 0.0, b = 0.0;
  for (int t = sl; t < T; t += 4) {
    const double x = in[(r0 + t) * kTailH + c];
    a += x;
    b += x * x;
  }
  s1[sl][c] = a;
  s2[sl][c] = b;
  __syncthreads();
  const int cpg = kTailH / groups;
  if (static_cast<int>(threadIdx.x) < groups) {
    const int g = threadIdx.x;
    double sa = 0.0, sb = 0.0;
    for (int k = 0; k < cpg; ++k)
      for (int q = 0; q < 4; ++q) { sa += s1[q][g * cpg + k]; sb += s2[q][g * cpg + k]; }
    const double n = static_cast<double>(T) * cpg;
    const double mean = n > 0 ? sa / n : 0.0;
    double var = n > 0 ? sb / n - mean * mean : 0.0;
    if (var < 0.0) var = 0.0;
    stats[(static_cast<int64_t>(v) * groups + g) * 2 + 0] = static_cast<float>(mean);
    stats[(static_cast<int64_t>(v) * groups + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

struct TailTile {  // the 16-byte tile record of convgemm::Tile plus the video index
  long long row0;
  int t0;
  int T;
};

constexpr int kTailRows = 128;
constexpr int kTailLd = kTailRows + 4;  // row pitch of the TRANSPOSED tile Xs[k][row]: 16-byte aligned row quads, and
                                        // 132 = 4 (mod 32) makes the transposing stores below conflict-free
constexpr int kTailThreads = 256;
// 256 threads: thread (rg, cg) owns rows rg*4 .. +3 and classes cg*CG .. +CG-1 (8 threads share a row quad: covers
// num_classes <= 8 * CG).  The activation tile is kept transposed in shared memory, so one LDS.128 fetches the four
// rows' values of a k and the inner loop is 1 + CG/2 shared loads per 4*CG FFMAs; two CTAs (16 warps) per SM.
template <int CG>
__global__ void __launch_bounds__(kTailThreads, 2)
tail_cls_lsm_kernel(const float* __restrict__ x, const float* __restrict__ stats, const int32_t* __restrict__ tile_vid,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ Wc /*[H][NC]*/,
                    const float* __restrict__ bc, const TailTile* __restrict__ tiles, int num_tiles, int groups, int relu,
                    int NC, float* __restrict__ z_out, float* __restrict__ lsm_out) {
  extern __shared__ __align__(16) float tsm[];
  __shared__ __align__(16) float gn_s[4][kTailH];   // mean, rstd (of the channel's group), gamma, beta of the tile's video
  float* Xs = tsm;                               // [128 k][132] (transposed: row index contiguous)
  float* Ws = tsm + kTailH * kTailLd;            // [128 k][8 * CG] zero padded
  float* bs = Ws + kTailH * 8 * CG;              // [8 * CG]
  constexpr int NCP = 8 * CG;
  for (int i = threadIdx.x; i < kTailH * NCP; i += blockDim.x) {
    const int k = i / NCP, n = i - k * NCP;
    Ws[i] = n < NC ? Wc[k * NC + n] : 0.f;
  }
  for (int i = threadIdx.x; i < NCP; i += blockDim.x) bs[i] = i < NC ? bc[i] : 0.f;
  const int cpg = kTailH / groups;
  const int rg = threadIdx.x >> 3, cg = threadIdx.x & 7;
  for (int ti = blockIdx.x; ti < num_tiles; ti += gridDim.x) {
    const TailTile tl = tiles[ti];
    const int v = tile_vid[ti];
    const int nrow = min(kTailRows, tl.T - tl.t0);
    const int64_t rbase = tl.row0 + tl.t0;
    __syncthreads();  // Xs of the previous tile is no longer read (and Ws / bs are written)
    // the tile's video: per-channel GroupNorm constants once per tile (the same four operands the element formula
    // below uses, so the arithmetic -- and the log-probabilities -- stay bit-identical to groupnorm_kernel's)
    if (threadIdx.x < kTailH) {
      const int c = threadIdx.x;
      const float* st = stats + (static_cast<int64_t>(v) * groups + c / cpg) * 2;
      gn_s[0][c] = st[0];
      gn_s[1][c] = st[1];
      gn_s[2][c] = __ldg(gamma + c);
      gn_s[3][c] = __ldg(beta + c);
    }
    __syncthreads();
    // GroupNorm + ReLU of the tile into shared memory (and to z_out).  A warp instruction covers 16 rows x 2 channel
    // quads: every row is read as one full 32-byte sector, and the four transposed stores of a lane hit banks
    // (quad*16 + e*4 + row) mod 32: all different across the warp.
    // (the 16 items of a thread in two batches of eight loads issued before any is used: ncu showed 65 % of the
    // kernel's stalls waiting on these loads when each iteration loaded, computed and stored in turn)
    constexpr int kItems = kTailRows * (kTailH / 4) / kTailThreads;   // 16
    constexpr int kBatch = 8;
#pragma unroll
    for (int i0 = 0; i0 < kItems; i0 += kBatch) {
      float4 xv[kBatch];
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const int i = threadIdx.x + (i0 + q) * kTailThreads;
        const int rsub = i & 15, cq = (i >> 4) & 1, rest = i >> 5;
        const int r = (rest & 7) * 16 + rsub, c4 = ((rest >> 3) * 2 + cq) * 4;
        xv[q] = r < nrow ? *reinterpret_cast<const float4*>(x + (rbase + r) * kTailH + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const int i = threadIdx.x + (i0 + q) * kTailThreads;
        const int rsub = i & 15, cq = (i >> 4) & 1, rest = i >> 5;
        const int r = (rest & 7) * 16 + rsub, c4 = ((rest >> 3) * 2 + cq) * 4;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrow) {
          const float4 mu = *reinterpret_cast<const float4*>(&gn_s[0][c4]), rs = *reinterpret_cast<const float4*>(&gn_s[1][c4]);
          const float4 ga = *reinterpret_cast<const float4*>(&gn_s[2][c4]), be = *reinterpret_cast<const float4*>(&gn_s[3][c4]);
          o.x = (xv[q].x - mu.x) * rs.x * ga.x + be.x;
          o.y = (xv[q].y - mu.y) * rs.y * ga.y + be.y;
          o.z = (xv[q].z - mu.z) * rs.z * ga.z + be.z;
          o.w = (xv[q].w - mu.w) * rs.w * ga.w + be.w;
          if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (z_out) *reinterpret_cast<float4*>(z_out + (rbase + r) * kTailH + c4) = o;
        }
        Xs[(c4 + 0) * kTailLd + r] = o.x;
        Xs[(c4 + 1) * kTailLd + r] = o.y;
        Xs[(c4 + 2) * kTailLd + r] = o.z;
        Xs[(c4 + 3) * kTailLd + r] = o.w;
      }
    }
    __syncthreads();
    // classifier: one accumulator per (row, class), k ascending, bias last (the order of conv1d_kernel).  For an even CG
    // two classes share one packed instruction (fma.rn.f32x2: two independent IEEE fp32 FMAs, same results as two
    // scalar FMAs, half the issue slots -- the kernel is issue-bound)
    float acc[4][CG];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CG; ++j) acc[i][j] = 0.f;
    const float* xr = Xs + rg * 4;
    const float* wr = Ws + cg * CG;
    if constexpr (CG % 2 == 0) {
      unsigned long long acc2[4][CG / 2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CG / 2; ++j) acc2[i][j] = 0ull;   // (+0.f, +0.f)
#pragma unroll 8
      for (int k = 0; k < kTailH; ++k) {
        const float4 x4 = *reinterpret_cast<const float4*>(xr + k * kTailLd);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
        unsigned long long w2[CG / 2];
#pragma unroll
        for (int j = 0; j < CG / 2; ++j) w2[j] = *reinterpret_cast<const unsigned long long*>(wr + k * NCP + 2 * j);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          unsigned long long xx;
          asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(xv[i]));
#pragma unroll
          for (int j = 0; j < CG / 2; ++j)
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i][j]) : "l"(xx), "l"(w2[j]));
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CG / 2; ++j)
          asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][2 * j]), "=f"(acc[i][2 * j + 1]) : "l"(acc2[i][j]));
    } else {
#pragma unroll 8
      for (int k = 0; k < kTailH; ++k) {
        const float4 x4 = *reinterpret_cast<const float4*>(xr + k * kTailLd);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
        float wv[CG];
#pragma unroll
        for (int j = 0; j < CG; ++j) wv[j] = wr[k * NCP + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < CG; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      }
    }
    // + bias, log-softmax over the row's classes (8 lanes x CG classes)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < CG; ++j) {
        acc[i][j] += bs[cg * CG + j];
        if (cg * CG + j < NC) m = fmaxf(m, acc[i][j]);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sden = 0.f;
#pragma unroll
      for (int j = 0; j < CG; ++j)
        if (cg * CG + j < NC) sden += expf(acc[i][j] - m);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) sden += __shfl_xor_sync(0xffffffffu, sden, o);
      const float lse = m + logf(sden);
      const int r = rg * 4 + i;
      if (r < nrow) {
#pragma unroll
        for (int j = 0; j < CG; ++j)
          if (cg * CG + j < NC) lsm_out[(rbase + r) * NC + cg * CG + j] = acc[i][j] - lse;
      }
    }
  }
}

// out[t, :] = table[min(floor(t * (float)Tz / T), Tz - 1), :]: the nearest-neighbour expansion alone
__global__ void __launch_bounds__(256) expand_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ off_z,
                                                          const int64_t* __restrict__ off_t, int C, float* __restrict__ out) {
  const int v = blockIdx.y;
  const int64_t z0 = off_z[v], t0v = off_t[v];
  const int Tz = static_cast<int>(off_z[v + 1] - z0);
  const int T = static_cast<int>(off_t[v + 1] - t0v);
  const float scale = static_cast<float>(Tz) / static_cast<float>(T);
  const int64_t n = static_cast<int64_t>(T) * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / C), c = static_cast<int>(i - static_cast<int64_t>(t) * C);
    int iz = static_cast<int>(floorf(static_cast<float>(t) * scale));
    iz = iz > Tz - 1 ? Tz - 1 : iz;
    out[t0v * C + i] = table[(z0 + iz) * C + c];
  }
}

}  // namespace
}  // namespace mucon

extern "C" int mucon_tail_logprobs(const float* x, const int64_t* row_off, const void* tiles, const int32_t* tile_vid,
                                   int num_tiles, int V, int H, int groups, float eps, int relu, const float* gamma,
                                   const float* beta, const float* Wc_hc, const float* bc, int num_classes,
                                   float* stats_ws, float* z_out, float* lsm_out, void* stream) {
  if (!x || !row_off || !tiles || !tile_vid || !gamma || !beta || !Wc_hc || !bc || !stats_ws || !lsm_out || V < 0 ||
      num_tiles < 0 || groups < 1 || num_classes < 1)
    return MUCON_EINVAL;
  if (H != kTailH || kTailH % groups != 0 || num_classes > 64) return MUCON_EUNSUPPORTED;
  if (V == 0 || num_tiles == 0) return MUCON_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  gn_stats_kernel<<<V, 512, 0, st>>>(x, row_off, groups, eps, stats_ws);
  MUCON_CUDA_CHECK(cudaGetLastError());
  const int CG = num_classes <= 24 ? 3 : (num_classes <= 48 ? 6 : 8);
  const size_t smem = sizeof(float) * (kTailH * kTailLd + kTailH * 8 * CG + 8 * CG);
  const int sms = mucon_device_sm_count();
  int grid = 2 * sms;
  if (grid > num_tiles) grid = num_tiles;
  const TailTile* tl = static_cast<const TailTile*>(tiles);
#define MUCON_TAIL(cg)                                                                                           \
  do {                                                                                                           \
    MUCON_CUDA_CHECK(cudaFuncSetAttribute(tail_cls_lsm_kernel<cg>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          static_cast<int>(smem)));                                              \
    tail_cls_lsm_kernel<cg><<<grid, kTailThreads, smem, st>>>(x, stats_ws, tile_vid, gamma, beta, Wc_hc, bc, tl, num_tiles, \
                                                     groups, relu, num_classes, z_out, lsm_out);                 \
  } while (0)
  if (CG == 3) MUCON_TAIL(3); else if (CG == 6) MUCON_TAIL(6); else MUCON_TAIL(8);
#undef MUCON_TAIL
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

// ---------------------------------------------------------------------------------------------
// training step: backward of the tail (autograd of models.py:759-768 GroupNorm + ReLU and :574-577 nearest upsample)
namespace mucon {
namespace {

// One CTA per video, 512 threads = 4 row slices x 128 channels.  y = relu(xhat * gamma + beta), xhat = (x - mean) * rstd:
//   g = dy * [y > 0];  dgamma[c] += sum_t g * xhat;  dbeta[c] += sum_t g;  dxhat = g * gamma
//   dx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat * xhat))      (means over the group's T x C/groups elements)
__global__ void __launch_bounds__(512) groupnorm_relu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 const int64_t* __restrict__ off, int groups, float eps,
                                                                 int relu, float* __restrict__ dx,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double s1[4][kTailH], s2[4][kTailH];
  __shared__ float mean_s[kTailH], rstd_s[kTailH], m1_s[kTailH], m2_s[kTailH];
  const int v = blockIdx.x;
  const int64_t r0 = off[v];
  const int T = static_cast<int>(off[v + 1] - r0);
  if (T == 0) return;
  const int c = threadIdx.x & (kTailH - 1), sl = threadIdx.x >> 7;
  const int cpg = kTailH / groups;
  const double n = static_cast<double>(T) * cpg;
  // pass 1: statistics (as gn_stats_kernel)
  double a = 0.0, b = 0.0;
  for (int t = sl; t < T; t += 4) {
    const double xv = x[(r0 + t) * kTailH + c];
    a += xv;
    b += xv * xv;
  }
  s1[sl][c] = a;
  s2[sl][c] = b;
  __syncthreads();
  if (threadIdx.x < kTailH) {
    const int g = c / cpg;
    double sa = 0.0, sb = 0.0;
    for (int k = 0; k < cpg; ++k)
      for (int q = 0; q < 4; ++q) { sa += s1[q][g * cpg + k]; sb += s2[q][g * cpg + k]; }
    const double mean = sa / n;
    double var = sb / n - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_s[c] = static_cast<float>(mean);
    rstd_s[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  // pass 2: per-channel sums of g and g * xhat
  const float mu = mean_s[c], rs = rstd_s[c], ga = gamma[c], be = beta[c];
  a = 0.0;
  b = 0.0;
  for (int t = sl; t < T; t += 4) {
    const float xh = (x[(r0 + t) * kTailH + c] - mu) * rs;
    float g = dy[(r0 + t) * kTailH + c];
    if (relu && xh * ga + be <= 0.f) g = 0.f;
    a += g;
    b += static_cast<double>(g) * xh;
  }
  s1[sl][c] = a;
  s2[sl][c] = b;
  __syncthreads();
  if (threadIdx.x < kTailH) {
    const double cg = s1[0][c] + s1[1][c] + s1[2][c] + s1[3][c];
    const double cgx = s2[0][c] + s2[1][c] + s2[2][c] + s2[3][c];
    atomicAdd(dbeta + c, static_cast<float>(cg));
    atomicAdd(dgamma + c, static_cast<float>(cgx));
    s1[0][c] = cg * ga;    // sum of dxhat over the channel
    s2[0][c] = cgx * ga;   // sum of dxhat * xhat over the channel
  }
  __syncthreads();
  if (threadIdx.x < kTailH) {
    const int g = c / cpg;
    double sa = 0.0, sb = 0.0;
    for (int k = 0; k < cpg; ++k) { sa += s1[0][g * cpg + k]; sb += s2[0][g * cpg + k]; }
    m1_s[c] = static_cast<float>(sa / n);
    m2_s[c] = static_cast<float>(sb / n);
  }
  __syncthreads();
  // pass 3: dx
  const float m1 = m1_s[c], m2 = m2_s[c];
  for (int t = sl; t < T; t += 4) {
    const float xh = (x[(r0 + t) * kTailH + c] - mu) * rs;
    float g = dy[(r0 + t) * kTailH + c];
    if (relu && xh * ga + be <= 0.f) g = 0.f;
    dx[(r0 + t) * kTailH + c] = rs * (g * ga - m1 - xh * m2);
  }
}

// grad_table[iz, :] = sum of grad_out[t, :] over the frames t whose nearest-neighbour source row is iz
// (iz(t) = min(floor(t * (float)Tz / T), Tz - 1) is non-decreasing in t: a contiguous range of frames per row)
__global__ void __launch_bounds__(256) expand_rows_bwd_kernel(const float* __restrict__ gout,
                                                              const int64_t* __restrict__ off_z,
                                                              const int64_t* __restrict__ off_t, int C,
                                                              float* __restrict__ gtable) {
  const int v = blockIdx.y;
  const int64_t z0 = off_z[v], t0v = off_t[v];
  const int Tz = static_cast<int>(off_z[v + 1] - z0);
  const int T = static_cast<int>(off_t[v + 1] - t0v);
  if (Tz == 0) return;
  const float scale = static_cast<float>(Tz) / static_cast<float>(T);
  auto src = [&](int t) {
    int iz = static_cast<int>(floorf(static_cast<float>(t) * scale));
    return iz > Tz - 1 ? Tz - 1 : iz;
  };
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int iz = wid; iz < Tz; iz += nw) {
    // first frame mapping to >= iz: start from the real-valued estimate and correct with the exact formula
    int lo = static_cast<int>(static_cast<double>(iz) * T / Tz) - 2;
    if (lo < 0) lo = 0;
    while (lo < T && src(lo) < iz) ++lo;
    while (lo > 0 && src(lo - 1) >= iz) --lo;
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
      for (int t = lo; t < T && src(t) == iz; ++t) acc += gout[(t0v + t) * C + c];
      gtable[(z0 + iz) * C + c] = acc;
    }
  }
}

}  // namespace
}  // namespace mucon

extern "C" int mucon_groupnorm_relu_bwd(const float* x, const float* dy, const float* gamma, const float* beta,
                                        const int64_t* row_off, int V, int C, int groups, float eps, int relu, float* dx,
                                        float* dgamma, float* dbeta, void* stream) {
  if (!x || !dy || !gamma || !beta || !row_off || !dx || !dgamma || !dbeta || V < 0 || groups < 1) return MUCON_EINVAL;
  if (C != kTailH || kTailH % groups != 0) return MUCON_EUNSUPPORTED;
  if (V == 0) return MUCON_OK;
  groupnorm_relu_bwd_kernel<<<V, 512, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, gamma, beta, row_off, groups, eps,
                                                                             relu, dx, dgamma, dbeta);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_expand_rows_bwd(const float* grad_out, const int64_t* off_z, const int64_t* off_t, int V, int max_Tz,
                                     int C, float* grad_table, void* stream) {
  if (!grad_out || !off_z || !off_t || !grad_table || V < 0 || C < 1 || max_Tz < 0) return MUCON_EINVAL;
  if (V == 0 || max_Tz == 0) return MUCON_OK;
  if (V > 65535) return MUCON_EUNSUPPORTED;
  int bx = (max_Tz + 7) / 8;
  if (bx > 64) bx = 64;
  expand_rows_bwd_kernel<<<dim3(bx, V), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, off_z, off_t, C, grad_table);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_expand_rows(const float* table, const int64_t* off_z, const int64_t* off_t, int V, int max_T, int C,
                                 float* out, void* stream) {
  if (!table || !off_z || !off_t || !out || V < 0 || C < 1 || max_T < 0) return MUCON_EINVAL;
  if (V == 0 || max_T == 0) return MUCON_OK;
  if (V > 65535) return MUCON_EUNSUPPORTED;
  int bx = static_cast<int>((static_cast<int64_t>(max_T) * C + 255) / 256);
  if (bx > 64) bx = 64;
  expand_rows_kernel<<<dim3(bx, V), 256, 0, static_cast<cudaStream_t>(stream)>>>(table, off_z, off_t, C, out);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

// log_softmax of every row, one warp per row: the same operations in the same order as the rows
// logsoftmax_expand_kernel prepares in shared memory (max, sum of expf, m + logf(s), x - lse)
__global__ void __launch_bounds__(256) logsoftmax_rows_kernel(const float* __restrict__ logits, int64_t rows, int C,
                                                              float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nw = (gridDim.x * static_cast<int64_t>(blockDim.x)) >> 5;
  for (int64_t r = wid; r < rows; r += nw) {
    const float* row = logits + r * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float lse = m + logf(s);
    for (int c = lane; c < C; c += 32) out[r * C + c] = row[c] - lse;
  }
}

extern "C" int mucon_logsoftmax_rows(const float* logits, int64_t rows, int C, float* out, void* stream) {
  if (!logits || !out || rows < 0 || C < 1) return MUCON_EINVAL;
  if (rows == 0) return MUCON_OK;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  logsoftmax_rows_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, rows, C, out);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_logsoftmax_expand(const float* logits, const int64_t* off_z, const int64_t* off_t, int V, int max_T,
                                       int C, float* out, void* stream) {
  if (!logits || !off_z || !off_t || !out || V < 0 || C < 1 || max_T < 0) return MUCON_EINVAL;
  if (V == 0 || max_T == 0) return MUCON_OK;
  if (V > 65535) return MUCON_EUNSUPPORTED;
  int bx = (max_T + kLseFrames - 1) / kLseFrames;
  if (bx > 64) bx = 64;
  const size_t smem = sizeof(float) * kLseMaxRows * C;
  if (smem > 48 * 1024) return MUCON_EUNSUPPORTED;
  logsoftmax_expand_kernel<<<dim3(bx, V), 256, smem, static_cast<cudaStream_t>(stream)>>>(logits, off_z, off_t, C, out);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

#ifdef MUCON_LAYER_TRACE
// developer builds only (not part of include/mucon_b200.h): the clock stamps of the last wavenet_layer_kernel launch
extern "C" int mucon_debug_layer_trace(long long* out_h) {
  return cudaMemcpyFromSymbol(out_h, mucon::layer::g_trace, sizeof(long long) * 32 * 128) == cudaSuccess ? 0 : -3;
}
#endif
