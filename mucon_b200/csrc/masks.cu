// masks.cu -- length -> soft mask generation for the mutual-consistency loss (sm_100a).
//
// Replaces reference src/mucon/masks.py:19-74 (create_masks): cumsum -> affine map ->
// affine_grid + bilinear grid_sample (zero padding) of a 100-tap template, fused into one pass
// that writes each [M, T] mask row with coalesced, aligned float4 stores (one CTA per row) and never
// materialises the grid.
//
// Per row i of a video with target size T (all float32, torch's op order):
//   pi = cumsum(L)[i] - L[i];  Ls = L[i]*(1+2*ov);  pi -= Ls*(ov/2)              masks.py:58-62
//   s  = T/Ls;  x = ((pi + Ls/2) - T/2) / (-(Ls/2))                              masks.py:102-120
//   g_t = (2t+1)/T - 1   (align_corners=0)   |  2t/(T-1) - 1   (align_corners=1) affine_grid
//   gx = g_t*s + x
//   u  = ((gx+1)*W - 1)/2 (align_corners=0)  |  (gx+1)/2*(W-1) (align_corners=1) grid_sample
//   out[i,t] = tmpl[floor u]*(1-frac) + tmpl[floor u + 1]*frac, taps outside [0,W) are 0
// Backward: d out/d u = tmpl[floor u + 1] - tmpl[floor u]; u depends on L through pi and Ls.
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <math.h>

#include "common.cuh"

namespace mucon {
namespace {

constexpr int kW = 100;  // TEMPLATE_WIDTH, masks.py:32

__constant__ float c_tmpl[3][kW];
std::atomic<int> g_tmpl_ready[64];   // 0 = not uploaded, 1 = uploaded (per device)
std::mutex g_tmpl_mutex;

void fill_templates(float (*t)[kW]) {
  for (int i = 0; i < kW; ++i) t[0][i] = 1.0f;  // box, masks.py:42-43
  // gaussian: scipy.signal.gaussian(M=100, std=20) = exp(-0.5*((n-(M-1)/2)/std)^2), masks.py:34-41
  for (int i = 0; i < kW; ++i) {
    const double n = i - (kW - 1) / 2.0;
    t[1][i] = (float)exp(-0.5 * (n / 20.0) * (n / 20.0));
  }
  // trapezoid: 25-tap ramps 0.5 -> 1 and 1 -> 0.5 (torch.arange with step 0.02), masks.py:44-54
  for (int i = 0; i < kW; ++i) t[2][i] = 1.0f;
  for (int i = 0; i < 25; ++i) {
    t[2][i] = (float)(0.5 + 0.02 * i);
    t[2][kW - 25 + i] = (float)(1.0 + (-0.02) * i);
  }
}

int ensure_templates() {
  int dev = 0;
  MUCON_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && g_tmpl_ready[dev].load(std::memory_order_acquire)) return MUCON_OK;
  std::lock_guard<std::mutex> lock(g_tmpl_mutex);   // two host threads calling for the first time: one upload
  if (dev < 64 && g_tmpl_ready[dev].load(std::memory_order_acquire)) return MUCON_OK;
  float h[3][kW];
  fill_templates(h);
  MUCON_CUDA_CHECK(cudaMemcpyToSymbol(c_tmpl, h, sizeof(h)));
  if (dev < 64) g_tmpl_ready[dev].store(1, std::memory_order_release);
  return MUCON_OK;
}

struct RowGeom {
  float s, x, Ls, pi;
};

__device__ __forceinline__ int find_video(const int32_t* n_off, int V, int row) {
  int lo = 0, hi = V;  // largest v with n_off[v] <= row
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (n_off[mid] <= row) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ RowGeom geom_from(float cum, float Li, int T, float overlap) {
  RowGeom g;
  float pi = cum - Li;
  g.Ls = Li * (1.0f + 2 * overlap);
  pi = pi - g.Ls * (overlap / 2);
  g.pi = pi;
  const float Tf = static_cast<float>(T);
  g.s = Tf / g.Ls;
  g.x = ((pi + g.Ls / 2) - Tf / 2) / (-(g.Ls / 2));
  return g;
}

__device__ __forceinline__ RowGeom row_geom(const float* L, int r0, int i, int T, float overlap) {
  float cum = 0.f;
  for (int q = 0; q <= i; ++q) cum = cum + L[r0 + q];  // sequential, like torch.cumsum on 1-D
  return geom_from(cum, L[r0 + i], T, overlap);
}

// Row context of a 64-thread group.  Every thread reads the (broadcast) row -> video entry and the
// video's sizes; the lengths in front of the row are fetched by one thread each into shared memory and
// summed by the group's first thread in order (torch.cumsum on a 1-D tensor is sequential), so the
// prologue is two dependent global loads deep instead of one per segment.  *g is valid on return.
struct RowCtx {
  int T, v, i;
  long long base;
};
constexpr int kGroup = 64;            // threads per mask row
constexpr int kGroupsPerCta = 4;      // rows a CTA works on at a time
__device__ __forceinline__ RowCtx row_ctx(const float* L, const int32_t* n_off, const int32_t* Tv,
                                          const int64_t* out_off, const int32_t* row_vid, int V, int row,
                                          float overlap, RowGeom* g, float* Lp, int gt, int bar_id) {
  const int v = row_vid ? row_vid[row] : find_video(n_off, V, row);
  const int T = Tv[v];
  const int r0 = n_off[v], i = row - r0;
  RowCtx c;
  c.T = T;
  c.v = v;
  c.i = i;
  c.base = out_off[v] + static_cast<long long>(i) * T;
  if (i < kGroup) {
    if (gt <= i) Lp[gt] = L[r0 + gt];
    named_bar_sync(bar_id, kGroup);
    if (gt == 0) {
      float cum = 0.f;
      for (int q = 0; q <= i; ++q) cum = cum + Lp[q];
      *g = geom_from(cum, Lp[i], T, overlap);
    }
  } else if (gt == 0) {
    *g = row_geom(L, r0, i, T, overlap);
  }
  named_bar_sync(bar_id, kGroup);
  return c;
}

__device__ __forceinline__ float coord_u(const RowGeom& g, int t, int T, int align) {
  const float Tf = static_cast<float>(T);
  float gt;
  if (align) gt = (T > 1) ? (2.0f * t) / (Tf - 1.0f) - 1.0f : -1.0f;
  else gt = (2.0f * t + 1.0f) / Tf - 1.0f;
  const float gx = gt * g.s + g.x;
  return align ? ((gx + 1.f) / 2.f) * (kW - 1) : ((gx + 1.f) * kW - 1.f) / 2.f;
}

// The template lives in shared memory, padded with one zero on each side (index i+1 holds tap i), so a
// per-thread index costs one bank access instead of a serialised constant-cache lookup.
constexpr int kWP = kW + 2;
__device__ __forceinline__ void load_template(float* tp, int tmpl) {
  for (int i = threadIdx.x; i < kWP; i += blockDim.x) tp[i] = (i >= 1 && i <= kW) ? c_tmpl[tmpl][i - 1] : 0.f;
}
__device__ __forceinline__ float tap(const float* tp, int i) { return tp[i + 1]; }  // valid for -1 <= i <= kW

__device__ __forceinline__ float sample(const float* tp, float u) {
  const float fl = floorf(u);
  // far outside the template: both taps are padding (also keeps the int conversion in range)
  if (!(fl >= -1.f && fl < (float)kW)) return 0.f;
  const int i0 = static_cast<int>(fl);
  const float w1 = u - fl, w0 = 1.f - w1;
  return tap(tp, i0) * w0 + tap(tp, i0 + 1) * w1;
}

// Cheap screen for the box template (the default, masks.py:42-43 / default.py:77): almost every frame
// of a row is either well inside the window (both taps are 1: the mask is 1 up to the last bit of
// w0 + w1, far below the stated tolerance) or well outside it (0).  u' is the closed form
// u = a_t * Wn / Ls - c evaluated with two FMAs; it differs from the reference-order evaluation by at
// most ~3e-3 (the rounding of pi ~ 1e4 times Wn/Ls), so frames within 0.05 of a template edge -- the
// ramps, where the value and the gradient live -- take the exact path.
struct Screen {
  float k, c0, c1;  // u' = (t * c1 + c0) * k - c
  float c;
};
__device__ __forceinline__ Screen make_screen(const RowGeom& g, int T, int align) {
  Screen sc;
  const float Tf = static_cast<float>(T);
  if (align) { sc.k = (float)(kW - 1) / g.Ls; sc.c1 = (T > 1) ? Tf / (Tf - 1.f) : 0.f; sc.c0 = -g.pi; sc.c = 0.f; }
  else { sc.k = (float)kW / g.Ls; sc.c1 = 1.f; sc.c0 = 0.5f - g.pi; sc.c = 0.5f; }
  return sc;
}
// 1: inside (mask 1, slope 0), 0: outside (mask 0, slope 0), 2: near an edge -> exact evaluation
__device__ __forceinline__ int screen_box(const Screen& sc, int t) {
  const float u = fmaf(fmaf(static_cast<float>(t), sc.c1, sc.c0), sc.k, -sc.c);
  if (u > 0.05f && u < (float)(kW - 1) - 0.05f) return 1;
  if (u < -1.05f || u > (float)kW + 0.05f) return 0;
  return 2;
}
// Frame ranges of a row: [0,a0) mask 0, [a0,a1) exact, [a1,b0) mask 1, [b0,b1) exact, [b1,T) mask 0.
// u' is non-decreasing in t, so each constant range is certified by screening its end points; the
// inverse map only proposes them.  Other templates: everything is "exact".
struct Regions {
  int a0, a1, b0, b1;
};
__device__ __forceinline__ Regions make_regions(const Screen& sc, int T, bool box) {
  Regions r;
  r.a0 = 0; r.a1 = T; r.b0 = T; r.b1 = T;
  if (!(sc.c1 > 0.f) || !(sc.k > 0.f) || T < 8) return r;
  auto inv = [&](float u) { return ((u + sc.c) / sc.k - sc.c0) / sc.c1; };
  auto clampi = [&](float x) { return x < 0.f ? 0 : (x > (float)T ? T : static_cast<int>(x)); };
  // every template is zero outside (-1, W): [0,a0) and [b1,T)
  int a0 = clampi(inv(-1.05f) - 1.f), b1 = clampi(inv((float)kW + 0.05f) + 2.f);
  for (int it = 0; it < 4 && a0 > 0 && screen_box(sc, a0 - 1) != 0; ++it) --a0;
  if (a0 > 0 && screen_box(sc, a0 - 1) != 0) a0 = 0;
  for (int it = 0; it < 4 && b1 < T && screen_box(sc, b1) != 0; ++it) ++b1;
  if (b1 < T && screen_box(sc, b1) != 0) b1 = T;
  if (b1 < a0) b1 = a0;
  r.a0 = a0; r.a1 = b1; r.b0 = b1; r.b1 = b1;  // one exact range [a0, b1)
  if (!box) return r;
  // the box template is 1 (slope 0) well inside the window: [a1,b0)
  int a1 = clampi(inv(0.05f) + 2.f), b0 = clampi(inv((float)(kW - 1) - 0.05f) - 1.f);
  if (a1 < a0) a1 = a0;
  if (b0 > b1) b0 = b1;
  if (a1 < b0) {
    for (int it = 0; it < 4 && a1 < b0 && screen_box(sc, a1) != 1; ++it) ++a1;
    for (int it = 0; it < 4 && b0 > a1 && screen_box(sc, b0 - 1) != 1; ++it) --b0;
    if (a1 < b0 && screen_box(sc, a1) == 1 && screen_box(sc, b0 - 1) == 1) { r.a1 = a1; r.b0 = b0; }
  }
  return r;
}
__device__ __forceinline__ float mask_value(const float* tp, const RowGeom& g, const Regions& r, int t, int T, int align) {
  if (t < r.a0 || t >= r.b1) return 0.f;
  if (t >= r.a1 && t < r.b0) return 1.f;
  return sample(tp, coord_u(g, t, T, align));
}

// Persistent CTAs of four 64-thread groups; a group takes one mask row at a time (grid-stride over the
// rows): a row is only ~2000 frames and starts with a chain of dependent loads, so what matters is how
// many rows are in flight per SM and that CTA launches are not the bottleneck (one CTA per row was
// bound by the block scheduler at ~280 CTAs/us).  Every thread writes 16-byte aligned float4 groups (rows
// start at arbitrary element offsets, so the first <= 3 and last <= 3 elements of a row are scalar stores).
__global__ void __launch_bounds__(kGroup * kGroupsPerCta)
masks_fwd_kernel(const float* __restrict__ L, const int32_t* __restrict__ n_off, const int32_t* __restrict__ Tv,
                 const int64_t* __restrict__ out_off, const int32_t* __restrict__ row_vid, int V, int n_rows,
                 float overlap, int tmpl, int align, float* __restrict__ L_scaled, float* __restrict__ out) {
  __shared__ float tp[kWP];
  __shared__ RowGeom g_s[kGroupsPerCta];
  __shared__ float Lp[kGroupsPerCta][kGroup];
  const int grp = threadIdx.x / kGroup, gt = threadIdx.x % kGroup;
  load_template(tp, tmpl);
  __syncthreads();
  const bool box = tmpl == 0;
  for (int row = blockIdx.x * kGroupsPerCta + grp; row < n_rows; row += gridDim.x * kGroupsPerCta) {
    const RowCtx c = row_ctx(L, n_off, Tv, out_off, row_vid, V, row, overlap, &g_s[grp], Lp[grp], gt, 1 + grp);
    const RowGeom g = g_s[grp];
    if (gt == 0 && L_scaled) L_scaled[row] = g.Ls;
    const int T = c.T;
    float* o = out + c.base;
    const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(o) >> 2) & 3);
    const int head = min(T, (4 - mis) & 3);
    const Regions r = make_regions(make_screen(g, T, align), T, box);
    if (gt < head) o[gt] = mask_value(tp, g, r, gt, T, align);
    const int nvec = (T - head) >> 2;
    float4* o4 = reinterpret_cast<float4*>(o + head);
    for (int q = gt; q < nvec; q += kGroup) {
      const int t = head + 4 * q;
      float4 val;
      if (t >= r.a1 && t + 3 < r.b0) val = make_float4(1.f, 1.f, 1.f, 1.f);
      else if (t + 3 < r.a0 || t >= r.b1) val = make_float4(0.f, 0.f, 0.f, 0.f);
      else val = make_float4(mask_value(tp, g, r, t, T, align), mask_value(tp, g, r, t + 1, T, align),
                             mask_value(tp, g, r, t + 2, T, align), mask_value(tp, g, r, t + 3, T, align));
      o4[q] = val;
    }
    const int t_tail = head + 4 * nvec + gt;
    if (t_tail < T) o[t_tail] = mask_value(tp, g, r, t_tail, T, align);
    named_bar_sync(1 + grp, kGroup);  // the group's shared slots are rewritten by the next row
  }
}

// Same decomposition: A = dLoss/dpi, B = dLoss/dLs of one row per 64-thread group.
__global__ void __launch_bounds__(kGroup * kGroupsPerCta)
masks_bwd_rows_kernel(const float* __restrict__ L, const int32_t* __restrict__ n_off, const int32_t* __restrict__ Tv,
                      const int64_t* __restrict__ out_off, const int32_t* __restrict__ row_vid, int V, int n_rows,
                      float overlap, int tmpl, int align, const float* __restrict__ gout, float* __restrict__ ws) {
  __shared__ float tp[kWP];
  __shared__ RowGeom g_s[kGroupsPerCta];
  __shared__ float Lp[kGroupsPerCta][kGroup];
  __shared__ float redA[kGroupsPerCta][kGroup / 32], redB[kGroupsPerCta][kGroup / 32];
  const int grp = threadIdx.x / kGroup, gt = threadIdx.x % kGroup;
  load_template(tp, tmpl);
  __syncthreads();
  const bool box = tmpl == 0;
  // u = a_t * Wn / Ls - c,  a_t = (t + 0.5 - pi) [align 0, Wn = W] or (t*T/(T-1) - pi) [align 1, Wn = W-1]
  const float Wn = align ? (float)(kW - 1) : (float)kW;
  for (int row = blockIdx.x * kGroupsPerCta + grp; row < n_rows; row += gridDim.x * kGroupsPerCta) {
    const RowCtx c = row_ctx(L, n_off, Tv, out_off, row_vid, V, row, overlap, &g_s[grp], Lp[grp], gt, 1 + grp);
    const RowGeom g = g_s[grp];
    const int T = c.T;
    const float* go = gout + c.base;
    const float Tf = (float)T;
    float A = 0.f, B = 0.f;
    const Regions r = make_regions(make_screen(g, T, align), T, box);
    // the box template has no slope away from its edges: only the two exact ranges contribute
    for (int pass = 0; pass < 2; ++pass) {
      const int lo = pass ? r.b0 : r.a0, hi = pass ? r.b1 : r.a1;
      for (int t = lo + gt; t < hi; t += kGroup) {
        const float u = coord_u(g, t, T, align);
        const float fl = floorf(u);
        if (!(fl >= -1.f && fl < (float)kW)) continue;
        const int i0 = static_cast<int>(fl);
        const float slope = tap(tp, i0 + 1) - tap(tp, i0);
        if (slope == 0.f) continue;
        const float at = align ? ((T > 1) ? (float)t * Tf / (Tf - 1.f) : 0.f) - g.pi : ((float)t + 0.5f) - g.pi;
        const float w = go[t] * slope;
        A += w * (-Wn / g.Ls);
        B += w * (-(at * Wn) / (g.Ls * g.Ls));
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      A += __shfl_xor_sync(0xffffffffu, A, off);
      B += __shfl_xor_sync(0xffffffffu, B, off);
    }
    if ((gt & 31) == 0) { redA[grp][gt >> 5] = A; redB[grp][gt >> 5] = B; }
    named_bar_sync(1 + grp, kGroup);
    if (gt == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < kGroup / 32; ++w) { a += redA[grp][w]; b += redB[grp][w]; }
      ws[2 * row] = a;
      ws[2 * row + 1] = b;
    }
    named_bar_sync(1 + grp, kGroup);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused "flint" evidence (models.py:456-468): E[r, c] = sum_t mask_r[t] * seg[t, c] without the masks
// ever being written: a row only reads the frames of its own window (4*T*C bytes per video in total).
// One 256-thread CTA per mask row at a time (grid-stride): a thread owns four class columns and every
// (256 / (C/4))-th frame of the row's window (float4 loads, several in flight), partial sums meet in
// shared memory.  Windows range from a few to thousands of frames, so the frames of one row are
// spread over the whole CTA rather than over a 64-thread group (the longest window was the tail).
constexpr int kFlintThreads = 256;
__global__ void __launch_bounds__(kFlintThreads)
flint_fwd_kernel(const float* __restrict__ L, const int32_t* __restrict__ n_off, const int32_t* __restrict__ Tv,
                 const int64_t* __restrict__ seg_off, const int32_t* __restrict__ row_vid, int V, int n_rows, int C,
                 float overlap, int tmpl, int align, const float* __restrict__ seg, float* __restrict__ E) {
  __shared__ float tp[kWP];
  __shared__ RowGeom g_s;
  __shared__ float Lp[kFlintThreads];
  __shared__ float4 part[kFlintThreads];
  const int tid = threadIdx.x;
  load_template(tp, tmpl);
  __syncthreads();
  const bool box = tmpl == 0;
  const int C4 = C >> 2;
  const int lanes_t = kFlintThreads / C4;
  const int tq = tid / C4, c4 = tid - tq * C4;
  const bool worker = tq < lanes_t;
  for (int row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int v = row_vid ? row_vid[row] : find_video(n_off, V, row);
    const int T = Tv[v];
    const int r0 = n_off[v], i = row - r0;
    if (i < kFlintThreads) {
      if (tid <= i) Lp[tid] = L[r0 + tid];
      __syncthreads();
      if (tid == 0) {
        float cum = 0.f;
        for (int q = 0; q <= i; ++q) cum = cum + Lp[q];
        g_s = geom_from(cum, Lp[i], T, overlap);
      }
    } else if (tid == 0) {
      g_s = row_geom(L, r0, i, T, overlap);
    }
    __syncthreads();
    const RowGeom g = g_s;
    const Regions r = make_regions(make_screen(g, T, align), T, box);
    const float4* sv = reinterpret_cast<const float4*>(seg + seg_off[v] * C);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (worker) {
#pragma unroll 4
      for (int t = r.a0 + tq; t < r.b1; t += lanes_t) {
        const float m = mask_value(tp, g, r, t, T, align);
        const float4 x = __ldg(sv + static_cast<long long>(t) * C4 + c4);
        acc.x = fmaf(m, x.x, acc.x); acc.y = fmaf(m, x.y, acc.y); acc.z = fmaf(m, x.z, acc.z); acc.w = fmaf(m, x.w, acc.w);
      }
    }
    part[tid] = acc;
    __syncthreads();
    if (tid < C4) {
      float4 sum = part[tid];
      for (int q = 1; q < lanes_t; ++q) {
        const float4 p = part[q * C4 + tid];
        sum.x += p.x; sum.y += p.y; sum.z += p.z; sum.w += p.w;
      }
      reinterpret_cast<float4*>(E + static_cast<long long>(row) * C)[tid] = sum;
    }
    __syncthreads();
  }
}

// The same evidence with a WARP per (mask row, part): a row's window is cut into kFlintParts equal parts, item
// row * kFlintParts + p is taken by one warp, grid-stride -- no block-level barrier anywhere, 64 warps per SM each
// with several 16-byte loads in flight, and a 5000-frame window is eight items instead of one CTA's tail.  A lane owns
// four class columns and every (32 / (C/4))-th frame of its part.  The partial sums of a row go to a workspace and a
// small second kernel adds them in the fixed order p = 0 .. parts-1, so the result does not depend on the order in which
// the warps ran (a first version let the warp that finished a row last do the sum: the fence and the counter round trip
// per item cost more than the extra launch, 0.209 vs 0.196 ms on the c2 split).
// Per-row constants of the evidence kernels, computed once per call by a thread per row (flint_rows_pre_kernel): the
// items of flint_fwd_warp_kernel then start from ONE 64-byte record instead of a chain of dependent loads (row -> video
// -> lengths) and ~200 instructions of geometry per lane.
struct __align__(16) RowPre {
  RowGeom g;        // 16 B
  Regions r;        // 16 B
  long long seg;    // first frame of the row's video in the packed logits (frames)
  int T, pad0;
  int pad1[4];
};
static_assert(sizeof(RowPre) == 64, "RowPre is one 64-byte record");

__global__ void __launch_bounds__(128)
flint_rows_pre_kernel(const float* __restrict__ L, const int32_t* __restrict__ n_off, const int32_t* __restrict__ Tv,
                      const int64_t* __restrict__ seg_off, const int32_t* __restrict__ row_vid, int V, int n_rows,
                      float overlap, int tmpl, int align, RowPre* __restrict__ pre) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int v = row_vid ? row_vid[row] : find_video(n_off, V, row);
  const int T = Tv[v];
  const int r0 = n_off[v], i = row - r0;
  RowPre p;
  p.g = row_geom(L, r0, i, T, overlap);
  p.r = make_regions(make_screen(p.g, T, align), T, tmpl == 0);
  p.seg = seg_off[v];
  p.T = T;
  p.pad0 = 0;
  p.pad1[0] = p.pad1[1] = p.pad1[2] = p.pad1[3] = 0;
  pre[row] = p;
}

// create_masks with the rows' geometry precomputed the same way: a 64-thread group starts a row from its 64-byte record --
// no dependent metadata loads, no serial prefix sum by the group's first thread, no barriers, no redundant region search.
__global__ void __launch_bounds__(128)
masks_rows_pre_kernel(const float* __restrict__ L, const int32_t* __restrict__ n_off, const int32_t* __restrict__ Tv,
                      const int64_t* __restrict__ out_off, const int32_t* __restrict__ row_vid, int V, int n_rows,
                      float overlap, int tmpl, int align, RowPre* __restrict__ pre, float* __restrict__ L_scaled) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int v = row_vid ? row_vid[row] : find_video(n_off, V, row);
  const int T = Tv[v];
  const int r0 = n_off[v], i = row - r0;
  RowPre p;
  p.g = row_geom(L, r0, i, T, overlap);
  p.r = make_regions(make_screen(p.g, T, align), T, tmpl == 0);
  p.seg = out_off[v] + static_cast<long long>(i) * T;
  p.T = T;
  p.pad0 = 0;
  p.pad1[0] = p.pad1[1] = p.pad1[2] = p.pad1[3] = 0;
  pre[row] = p;
  if (L_scaled) L_scaled[row] = p.g.Ls;
}

__global__ void __launch_bounds__(kGroup * kGroupsPerCta)
masks_fwd_pre_kernel(const RowPre* __restrict__ pre, int n_rows, int tmpl, int align, float* __restrict__ out) {
  __shared__ float tp[kWP];
  const int grp = threadIdx.x / kGroup, gt = threadIdx.x % kGroup;
  load_template(tp, tmpl);
  __syncthreads();
  for (int row = blockIdx.x * kGroupsPerCta + grp; row < n_rows; row += gridDim.x * kGroupsPerCta) {
    const RowPre p = pre[row];
    const RowGeom g = p.g;
    const Regions r = p.r;
    const int T = p.T;
    float* o = out + p.seg;
    const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(o) >> 2) & 3);
    const int head = min(T, (4 - mis) & 3);
    if (gt < head) o[gt] = mask_value(tp, g, r, gt, T, align);
    const int nvec = (T - head) >> 2;
    float4* o4 = reinterpret_cast<float4*>(o + head);
    for (int q = gt; q < nvec; q += kGroup) {
      const int t = head + 4 * q;
      float4 val;
      if (t >= r.a1 && t + 3 < r.b0) val = make_float4(1.f, 1.f, 1.f, 1.f);
      else if (t + 3 < r.a0 || t >= r.b1) val = make_float4(0.f, 0.f, 0.f, 0.f);
      else val = make_float4(mask_value(tp, g, r, t, T, align), mask_value(tp, g, r, t + 1, T, align),
                             mask_value(tp, g, r, t + 2, T, align), mask_value(tp, g, r, t + 3, T, align));
      o4[q] = val;
    }
    const int t_tail = head + 4 * nvec + gt;
    if (t_tail < T) o[t_tail] = mask_value(tp, g, r, t_tail, T, align);
  }
}

constexpr int kFlintPartsMax = 8;
constexpr int kFlintBatch = 8;
template <int kFlintParts>
__global__ void __launch_bounds__(256, 4)
flint_fwd_warp_kernel(const RowPre* __restrict__ pre, int n_rows, int C, int tmpl, int align,
                      const float* __restrict__ seg, float* __restrict__ part_ws) {
  __shared__ float tp[kWP];
  load_template(tp, tmpl);
  __syncthreads();
  const int C4 = C >> 2;
  const int FP = 32 / C4;                       // frames per warp iteration
  const int lane = threadIdx.x & 31;
  const int fq = lane / C4, c4 = lane - fq * C4;
  const bool active = fq < FP;
  const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int n_items = n_rows * kFlintParts;
  for (int item = warp_g; item < n_items; item += n_warps) {
    const int row = item / kFlintParts, p = item - row * kFlintParts;
    const RowPre rp = pre[row];                                 // one 64-byte record, the same for every lane
    const RowGeom g = rp.g;
    const Regions r = rp.r;
    const int T = rp.T;
    const int len = r.b1 - r.a0;
    const int t_lo = r.a0 + static_cast<int>(static_cast<long long>(len) * p / kFlintParts);
    const int t_hi = r.a0 + static_cast<int>(static_cast<long long>(len) * (p + 1) / kFlintParts);
    const float4* sv = reinterpret_cast<const float4*>(seg + rp.seg * C);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      // batches of kFlintBatch independent 16-byte loads issued before any of them is used (the mask evaluation has
      // branches the compiler will not move loads across: without the explicit batch one load per warp was in flight)
      for (int t = t_lo + fq; t < t_hi; t += kFlintBatch * FP) {
        float4 x[kFlintBatch];
#pragma unroll
        for (int k = 0; k < kFlintBatch; ++k) {
          const int tt = t + k * FP;
          x[k] = tt < t_hi ? __ldg(sv + static_cast<long long>(tt) * C4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < kFlintBatch; ++k) {
          const int tt = t + k * FP;
          if (tt < t_hi) {
            const float m = mask_value(tp, g, r, tt, T, align);
            acc.x = fmaf(m, x[k].x, acc.x); acc.y = fmaf(m, x[k].y, acc.y);
            acc.z = fmaf(m, x[k].z, acc.z); acc.w = fmaf(m, x[k].w, acc.w);
          }
        }
      }
    }
    // frames-of-the-iteration groups -> group 0 (fixed order: fq = 0, 1, ...)
    for (int k = 1; k < FP; ++k) {
      const int src = lane + k * C4;
      const float x = __shfl_sync(0xffffffffu, acc.x, src & 31), y = __shfl_sync(0xffffffffu, acc.y, src & 31);
      const float z = __shfl_sync(0xffffffffu, acc.z, src & 31), w = __shfl_sync(0xffffffffu, acc.w, src & 31);
      if (fq == 0) { acc.x += x; acc.y += y; acc.z += z; acc.w += w; }
    }
    float4* pw = reinterpret_cast<float4*>(part_ws + (static_cast<long long>(row) * kFlintParts + p) * C);
    if (fq == 0 && c4 < C4) pw[c4] = acc;
  }
}

// E[row, :] = the row's partial sums added in the fixed order p = 0 .. parts-1 (a thread per (row, four classes)): the
// result does not depend on the order in which the warps of flint_fwd_warp_kernel ran, and those warps neither fence nor
// touch a counter.
__global__ void __launch_bounds__(256)
flint_fwd_combine_kernel(const float* __restrict__ part_ws, int n_rows, int C, int parts, float* __restrict__ E) {
  const int C4 = C >> 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(n_rows) * C4) return;
  const long long row = i / C4;
  const int c4 = static_cast<int>(i - row * C4);
  const float4* pr = reinterpret_cast<const float4*>(part_ws + row * parts * C);
  float4 sum = pr[c4];
  for (int q = 1; q < parts; ++q) {
    const float4 a = pr[q * C4 + c4];
    sum.x += a.x; sum.y += a.y; sum.z += a.z; sum.w += a.w;
  }
  reinterpret_cast<float4*>(E + row * C)[c4] = sum;
}

// d seg[t, :] = sum_r mask_r[t] * gE[r, :] over the rows of the frame's video.  One CTA per chunk of
// kFlintChunk frames of one video (host-built chunk list {video, t0}); the video's row geometry is
// set up once per CTA, then a thread owns one (frame, 4 classes) item at a time: coalesced stores.
constexpr int kFlintChunk = 512;
constexpr int kFlintMaxRows = 64;
struct FlintChunk {
  int v, t0;
};
__global__ void __launch_bounds__(256)
flint_bwd_seg_kernel(const float* __restrict__ L, const int32_t* __restrict__ n_off, const int32_t* __restrict__ Tv,
                     const int64_t* __restrict__ seg_off, const FlintChunk* __restrict__ chunks, int C, float overlap,
                     int tmpl, int align, const float* __restrict__ gE, float* __restrict__ gseg) {
  __shared__ float tp[kWP];
  __shared__ RowGeom g_s[kFlintMaxRows];
  __shared__ Regions r_s[kFlintMaxRows];
  const FlintChunk ck = chunks[blockIdx.x];
  const int T = Tv[ck.v], r0 = n_off[ck.v], N = n_off[ck.v + 1] - r0;
  load_template(tp, tmpl);
  if (static_cast<int>(threadIdx.x) < N) {
    const RowGeom g = row_geom(L, r0, threadIdx.x, T, overlap);
    g_s[threadIdx.x] = g;
    r_s[threadIdx.x] = make_regions(make_screen(g, T, align), T, tmpl == 0);
  }
  __syncthreads();
  const int C4 = C >> 2;
  const int nt = min(kFlintChunk, T - ck.t0);
  float4* out = reinterpret_cast<float4*>(gseg + (seg_off[ck.v] + ck.t0) * C);
  const float4* g4 = reinterpret_cast<const float4*>(gE + static_cast<long long>(r0) * C);
  for (int idx = threadIdx.x; idx < nt * C4; idx += blockDim.x) {
    const int tl = idx / C4, c4 = idx - tl * C4;
    const int t = ck.t0 + tl;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < N; ++r) {
      const Regions rg = r_s[r];
      if (t < rg.a0 || t >= rg.b1) continue;
      const float m = mask_value(tp, g_s[r], rg, t, T, align);
      const float4 gv = __ldg(g4 + r * C4 + c4);
      acc.x = fmaf(m, gv.x, acc.x); acc.y = fmaf(m, gv.y, acc.y); acc.z = fmaf(m, gv.z, acc.z); acc.w = fmaf(m, gv.w, acc.w);
    }
    out[idx] = acc;
  }
}

// dLoss/dpi and dLoss/dLs of one row per 64-thread group, with d mask[t] = sum_c gE[r, c] * seg[t, c]
// formed on the fly for the frames that have a slope (the ramps of the box template).
__global__ void __launch_bounds__(kGroup * kGroupsPerCta)
flint_bwd_rows_kernel(const float* __restrict__ L, const int32_t* __restrict__ n_off, const int32_t* __restrict__ Tv,
                      const int64_t* __restrict__ seg_off, const int32_t* __restrict__ row_vid, int V, int n_rows,
                      int C, float overlap, int tmpl, int align, const float* __restrict__ seg,
                      const float* __restrict__ gE, float* __restrict__ ws) {
  __shared__ float tp[kWP];
  __shared__ RowGeom g_s[kGroupsPerCta];
  __shared__ float Lp[kGroupsPerCta][kGroup];
  __shared__ float red[kGroupsPerCta][kGroup / 32];
  const int grp = threadIdx.x / kGroup, gt = threadIdx.x % kGroup;
  load_template(tp, tmpl);
  __syncthreads();
  const bool box = tmpl == 0;
  const float Wn = align ? (float)(kW - 1) : (float)kW;
  for (int row = blockIdx.x * kGroupsPerCta + grp; row < n_rows; row += gridDim.x * kGroupsPerCta) {
    const RowCtx c = row_ctx(L, n_off, Tv, seg_off, row_vid, V, row, overlap, &g_s[grp], Lp[grp], gt, 1 + grp);
    const RowGeom g = g_s[grp];
    const int T = c.T;
    const float Tf = (float)T;
    const Regions r = make_regions(make_screen(g, T, align), T, box);
    const float* sv = seg + seg_off[c.v] * C;
    const float ge0 = gt < C ? gE[static_cast<long long>(row) * C + gt] : 0.f;
    const float ge1 = gt + kGroup < C ? gE[static_cast<long long>(row) * C + gt + kGroup] : 0.f;
    float A = 0.f, B = 0.f;  // accumulated by the group's first thread
    for (int pass = 0; pass < 2; ++pass) {
      const int lo = pass ? r.b0 : r.a0, hi = pass ? r.b1 : r.a1;
      for (int t = lo; t < hi; ++t) {
        const float u = coord_u(g, t, T, align);
        const float fl = floorf(u);
        if (!(fl >= -1.f && fl < (float)kW)) continue;  // uniform over the group
        const int i0 = static_cast<int>(fl);
        const float slope = tap(tp, i0 + 1) - tap(tp, i0);
        if (slope == 0.f) continue;
        const float* sr = sv + static_cast<long long>(t) * C;
        float d = 0.f;
        if (gt < C) d = ge0 * sr[gt];
        if (gt + kGroup < C) d = fmaf(ge1, sr[gt + kGroup], d);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
        if ((gt & 31) == 0) red[grp][gt >> 5] = d;
        named_bar_sync(1 + grp, kGroup);
        if (gt == 0) {
          float dm = 0.f;
          for (int w = 0; w < kGroup / 32; ++w) dm += red[grp][w];
          const float at = align ? ((T > 1) ? (float)t * Tf / (Tf - 1.f) : 0.f) - g.pi : ((float)t + 0.5f) - g.pi;
          const float w = dm * slope;
          A += w * (-Wn / g.Ls);
          B += w * (-(at * Wn) / (g.Ls * g.Ls));
        }
        named_bar_sync(1 + grp, kGroup);
      }
    }
    if (gt == 0) { ws[2 * row] = A; ws[2 * row + 1] = B; }
    named_bar_sync(1 + grp, kGroup);
  }
}

// grad_L[j] = sum_{i>j} A_i - A_j*(1+2ov)*(ov/2) + B_j*(1+2ov)
__global__ void masks_bwd_combine_kernel(const int32_t* __restrict__ n_off, int V, float overlap,
                                         const float* __restrict__ ws, float* __restrict__ grad_L) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int r0 = n_off[v], r1 = n_off[v + 1];
  const float k = 1.0f + 2 * overlap;
  float suffix = 0.f;
  for (int r = r1 - 1; r >= r0; --r) {
    const float A = ws[2 * r], B = ws[2 * r + 1];
    grad_L[r] = suffix - A * (k * (overlap / 2)) + B * k;
    suffix += A;
  }
}

}  // namespace
}  // namespace mucon

using namespace mucon;

static int mask_grid(int n_rows) {
  const int sms = mucon_device_sm_count();   // per current device (cached per device in api.cu)
  const int want = (n_rows + kGroupsPerCta - 1) / kGroupsPerCta;
  const int cap = (sms > 0 ? sms : 148) * 8;  // 8 CTAs x 4 groups x 64 threads = 2048 threads per SM
  return want < cap ? want : cap;
}

extern "C" int mucon_masks_fwd(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* out_off,
                               const int32_t* row_vid, int V, int n_rows, int max_T, float overlap, int template_id,
                               int align_corners, float* L_scaled, float* out, void* stream) {
  if (!L || !n_off || !T || !out_off || !out || V < 0 || n_rows < 0 || max_T < 0) return MUCON_EINVAL;
  if (template_id < 0 || template_id > 2) return MUCON_EINVAL;
  if (V == 0 || n_rows == 0 || max_T == 0) return MUCON_OK;
  int rc = ensure_templates();
  if (rc != MUCON_OK) return rc;
  masks_fwd_kernel<<<mask_grid(n_rows), kGroup * kGroupsPerCta, 0, static_cast<cudaStream_t>(stream)>>>(L, n_off, T, out_off, row_vid, V, n_rows, overlap, template_id,
                                                                         align_corners, L_scaled, out);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

// ws: 64 bytes per mask row (16-byte aligned), overwritten
extern "C" int mucon_masks_fwd_ws(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* out_off,
                                  const int32_t* row_vid, int V, int n_rows, int max_T, float overlap, int template_id,
                                  int align_corners, float* L_scaled, float* out, void* ws, void* stream) {
  if (!L || !n_off || !T || !out_off || !out || !ws || V < 0 || n_rows < 0 || max_T < 0) return MUCON_EINVAL;
  if (template_id < 0 || template_id > 2) return MUCON_EINVAL;
  if (reinterpret_cast<uintptr_t>(ws) & 15) return MUCON_EALIGN;
  if (V == 0 || n_rows == 0 || max_T == 0) return MUCON_OK;
  int rc = ensure_templates();
  if (rc != MUCON_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RowPre* pre = static_cast<RowPre*>(ws);
  masks_rows_pre_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(L, n_off, T, out_off, row_vid, V, n_rows, overlap, template_id,
                                                              align_corners, pre, L_scaled);
  MUCON_CUDA_CHECK(cudaGetLastError());
  masks_fwd_pre_kernel<<<mask_grid(n_rows), kGroup * kGroupsPerCta, 0, st>>>(pre, n_rows, template_id, align_corners, out);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_masks_bwd(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* out_off,
                               const int32_t* row_vid, int V, int n_rows, float overlap, int template_id,
                               int align_corners, const float* grad_out, float* ws, float* grad_L, void* stream) {
  if (!L || !n_off || !T || !out_off || !grad_out || !ws || !grad_L || V < 0 || n_rows < 0) return MUCON_EINVAL;
  if (template_id < 0 || template_id > 2) return MUCON_EINVAL;
  if (V == 0 || n_rows == 0) return MUCON_OK;
  int rc = ensure_templates();
  if (rc != MUCON_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  masks_bwd_rows_kernel<<<mask_grid(n_rows), kGroup * kGroupsPerCta, 0, st>>>(L, n_off, T, out_off, row_vid, V, n_rows, overlap, template_id, align_corners, grad_out,
                                                ws);
  MUCON_CUDA_CHECK(cudaGetLastError());
  masks_bwd_combine_kernel<<<(V + 127) / 128, 128, 0, st>>>(n_off, V, overlap, ws, grad_L);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_mask_template_h(int template_id, float* out100_h) {
  if (!out100_h || template_id < 0 || template_id > 2) return MUCON_EINVAL;
  float h[3][kW];
  fill_templates(h);
  for (int i = 0; i < kW; ++i) out100_h[i] = h[template_id][i];
  return MUCON_OK;
}

extern "C" int mucon_flint_fwd(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* seg_off,
                               const int32_t* row_vid, int V, int n_rows, int C, float overlap, int template_id,
                               int align_corners, const float* seg, float* E, void* stream) {
  if (!L || !n_off || !T || !seg_off || !seg || !E || V < 0 || n_rows < 0 || C < 1) return MUCON_EINVAL;
  if (template_id < 0 || template_id > 2) return MUCON_EINVAL;
  if (C > 2 * kGroup || (C & 3)) return MUCON_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(seg) & 15) || (reinterpret_cast<uintptr_t>(E) & 15)) return MUCON_EALIGN;
  if (V == 0 || n_rows == 0) return MUCON_OK;
  int rc = ensure_templates();
  if (rc != MUCON_OK) return rc;
  const int sms = mucon_device_sm_count();   // per current device (cached per device in api.cu)
  const int grid = n_rows < 8 * sms ? n_rows : 8 * sms;
  flint_fwd_kernel<<<grid, kFlintThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      L, n_off, T, seg_off, row_vid, V, n_rows, C, overlap, template_id, align_corners, seg, E);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int64_t mucon_flint_fwd_ws_words(int n_rows, int C) {
  if (n_rows < 0 || C < 1) return 0;
  // float partials + one counter per row (+ padding to 16 bytes) + one 64-byte RowPre per row
  return static_cast<int64_t>(n_rows) * kFlintPartsMax * C + (static_cast<int64_t>(n_rows) + 3) / 4 * 4 +
         static_cast<int64_t>(n_rows) * 16;
}

// ws: mucon_flint_fwd_ws_words(n_rows, C) 4-byte words, 16-byte aligned: [partial sums | (unused) | row records]; overwritten.
extern "C" int mucon_flint_fwd_ws(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* seg_off,
                                  const int32_t* row_vid, int V, int n_rows, int C, float overlap, int template_id,
                                  int align_corners, const float* seg, float* ws, float* E, void* stream) {
  if (!L || !n_off || !T || !seg_off || !seg || !E || !ws || V < 0 || n_rows < 0 || C < 1) return MUCON_EINVAL;
  if (template_id < 0 || template_id > 2) return MUCON_EINVAL;
  if (C > 2 * kGroup || (C & 3)) return MUCON_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(seg) & 15) || (reinterpret_cast<uintptr_t>(E) & 15) ||
      (reinterpret_cast<uintptr_t>(ws) & 15))
    return MUCON_EALIGN;
  if (V == 0 || n_rows == 0) return MUCON_OK;
  int rc = ensure_templates();
  if (rc != MUCON_OK) return rc;
  const int sms = mucon_device_sm_count();
  static const int parts = getenv("MUCON_FLINT_PARTS") ? atoi(getenv("MUCON_FLINT_PARTS")) : 8;
  const long long items = static_cast<long long>(n_rows) * parts;
  static const int bps = getenv("MUCON_FLINT_BPS") ? atoi(getenv("MUCON_FLINT_BPS")) : 48;  // CTAs per SM in the grid
  // (measured on c2, 11839 rows: parts 4 / 8 CTAs per SM 0.196 ms; parts 8 / 24 per SM 0.174 ms -- finer items balance the
  // ragged windows and shorten the last wave; with the kernel held to 64 registers (four resident CTAs instead of three,
  // __launch_bounds__(256, 4)) and 48 per SM 0.160 ms; 48 registers spill and lose: 0.194 ms)
  long long grid = (items + 7) / 8;
  if (grid > static_cast<long long>(bps) * sms) grid = static_cast<long long>(bps) * sms;
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws + static_cast<int64_t>(n_rows) * kFlintPartsMax * C);
  RowPre* pre = reinterpret_cast<RowPre*>(counters + (static_cast<int64_t>(n_rows) + 3) / 4 * 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  flint_rows_pre_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(L, n_off, T, seg_off, row_vid, V, n_rows, overlap, template_id,
                                                              align_corners, pre);
  MUCON_CUDA_CHECK(cudaGetLastError());
  const int np = parts == 8 ? 8 : (parts == 2 ? 2 : 4);
  if (np == 8)
    flint_fwd_warp_kernel<8><<<static_cast<int>(grid), 256, 0, st>>>(pre, n_rows, C, template_id, align_corners, seg, ws);
  else if (np == 2)
    flint_fwd_warp_kernel<2><<<static_cast<int>(grid), 256, 0, st>>>(pre, n_rows, C, template_id, align_corners, seg, ws);
  else
    flint_fwd_warp_kernel<4><<<static_cast<int>(grid), 256, 0, st>>>(pre, n_rows, C, template_id, align_corners, seg, ws);
  MUCON_CUDA_CHECK(cudaGetLastError());
  const long long n_out = static_cast<long long>(n_rows) * (C >> 2);
  flint_fwd_combine_kernel<<<static_cast<int>((n_out + 255) / 256), 256, 0, st>>>(ws, n_rows, C, np, E);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_flint_bwd(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* seg_off,
                               const int32_t* row_vid, int V, int n_rows, int max_rows, int C, float overlap,
                               int template_id, int align_corners, const float* seg, const float* grad_E,
                               const void* chunks, int n_chunks, float* grad_seg, float* ws, float* grad_L,
                               void* stream) {
  if (!L || !n_off || !T || !seg_off || !seg || !grad_E || !ws || !grad_L || V < 0 || n_rows < 0 || C < 1)
    return MUCON_EINVAL;
  if (template_id < 0 || template_id > 2) return MUCON_EINVAL;
  if (grad_seg && (!chunks || n_chunks < 0)) return MUCON_EINVAL;
  if (C > 2 * kGroup || (C & 3) || max_rows > kFlintMaxRows) return MUCON_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(grad_E) & 15) || (grad_seg && (reinterpret_cast<uintptr_t>(grad_seg) & 15))) return MUCON_EALIGN;
  if (V == 0 || n_rows == 0) return MUCON_OK;
  int rc = ensure_templates();
  if (rc != MUCON_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grad_seg && n_chunks > 0) {
    flint_bwd_seg_kernel<<<n_chunks, 256, 0, st>>>(L, n_off, T, seg_off, static_cast<const FlintChunk*>(chunks), C,
                                                   overlap, template_id, align_corners, grad_E, grad_seg);
    MUCON_CUDA_CHECK(cudaGetLastError());
  }
  flint_bwd_rows_kernel<<<mask_grid(n_rows), kGroup * kGroupsPerCta, 0, st>>>(
      L, n_off, T, seg_off, row_vid, V, n_rows, C, overlap, template_id, align_corners, seg, grad_E, ws);
  MUCON_CUDA_CHECK(cudaGetLastError());
  masks_bwd_combine_kernel<<<(V + 127) / 128, 128, 0, st>>>(n_off, V, overlap, ws, grad_L);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}
