// shead.cu -- the sequence-generation ("s") head of the reference at test time (sm_100a).
//
// Replaces, for inference, reference src/mucon/models.py:585-745 (`sequence_generation_forward` and
// `_calculate_attention`): a bidirectional LSTM encoder over the encoded sequence z [Tz, 128], then an attention
// decoder (embedding -> additive attention over the encoder outputs -> attn_combine -> LSTM cell -> transcript head
// and length head) run step by step, teacher-forced or greedy.  The reference runs it one video at a time from Python
// with a device->host `.item()` per decoding step (models.py:721); here a whole batch of variable-length videos is
// two launches and the greedy loop (argmax, EOS test, next input) never leaves the GPU.
//
//   lstm_recurrent_kernel   one CTA per (four videos of similar length, direction): the recurrence
//                           h_t = LSTM(Xproj[t] + W_hh h_{t-1}); the input projections of ALL steps are one conv GEMM
//                           launched beforehand (mucon_conv1d).  W_hh (512 x 128 fp32 = 256 KB) lives half in registers
//                           (64 per thread, thread j owns gate row j) and half in shared memory; every weight fetched
//                           feeds the four recurrences; a step is 4 x 128 FMAs per thread and two barriers.
//   seq_decoder_kernel      one CTA per kDecB videos (= 1; more run in lockstep with finished ones masked): all decoding
//                           steps; every mat-vec is a warp per EIGHT output rows with the lanes across the input
//                           (coalesced weight reads from L2, eight independent rows in flight), fp32 throughout.
// Arithmetic is fp32 with expf / tanhf / logf (no fast-math); sums run in a different order than torch's LSTM /
// Linear kernels, so outputs agree with the reference to ~1e-5, not bit for bit (tests/test_shead.py).
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace mucon {
namespace {

constexpr int kH = 128;        // encoder hidden = decoder hidden = ft hidden (src/configs/mucon/default.py:98-116)
constexpr int kG = 4 * kH;     // LSTM gate rows, torch order i, f, g, o

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int kEncB = 4;   // videos per CTA: every weight fetched feeds four recurrences (8: 5.5 ms on c2, 4: 3.6 ms, 1: 6.1 ms)
constexpr int kEncP = kEncB * kH / kG;   // (video, unit) pairs per thread in the gate phase

__global__ void __launch_bounds__(kG, 1)
lstm_recurrent_kernel(const float* __restrict__ xproj_f, const float* __restrict__ xproj_b,
                      const float* __restrict__ whh_f, const float* __restrict__ whh_b,
                      const int64_t* __restrict__ row_off, const int32_t* __restrict__ order, int V,
                      float* __restrict__ enc_out /*[rows, 2H]*/, float* __restrict__ hn /*[V, 2, H]*/,
                      float* __restrict__ cn /*[V, 2, H]*/) {
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                    // [64][512]: W_hh[j][64 + k] at Ws[k * 512 + j]
  float* h_s = sm + 64 * kG;         // [kEncB][128]
  float* g_s = h_s + kEncB * kH;     // [kEncB][512]
  const int dir = blockIdx.y, j = threadIdx.x;
  const float* xproj = dir ? xproj_b : xproj_f;
  const float* whh = dir ? whh_b : whh_f;
  // the CTA's videos: kEncB consecutive entries of `order` (videos sorted by length, so a group's lengths are alike)
  int64_t r0[kEncB];
  int Tz[kEncB], vid[kEncB], Tmax = 0;
#pragma unroll
  for (int b = 0; b < kEncB; ++b) {
    const int idx = blockIdx.x * kEncB + b;
    vid[b] = idx < V ? order[idx] : -1;
    r0[b] = vid[b] >= 0 ? row_off[vid[b]] : 0;
    Tz[b] = vid[b] >= 0 ? static_cast<int>(row_off[vid[b] + 1] - r0[b]) : 0;
    Tmax = max(Tmax, Tz[b]);
  }
  float w[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) w[k] = whh[j * kH + k];
  for (int k = 0; k < 64; ++k) Ws[k * kG + j] = whh[j * kH + 64 + k];
  for (int i = j; i < kEncB * kH; i += kG) h_s[i] = 0.f;
  // phase-2 role of this thread: unit u of video slots (j >> 7) + 4p
  const int u = j & (kH - 1);
  float c[kEncP], h[kEncP];
#pragma unroll
  for (int p = 0; p < kEncP; ++p) c[p] = h[p] = 0.f;
  __syncthreads();
  // the input projections of step s + 1 are fetched while step s computes (an L2 round trip per step otherwise)
  float xnext[kEncB];
#pragma unroll
  for (int b = 0; b < kEncB; ++b) xnext[b] = Tz[b] > 0 ? xproj[(r0[b] + (dir ? Tz[b] - 1 : 0)) * kG + j] : 0.f;
  for (int s = 0; s < Tmax; ++s) {
    float acc[kEncB];
#pragma unroll
    for (int b = 0; b < kEncB; ++b) {
      acc[b] = xnext[b];
      if (s + 1 < Tz[b]) xnext[b] = xproj[(r0[b] + (dir ? Tz[b] - 2 - s : s + 1)) * kG + j];
    }
    const float4* h4 = reinterpret_cast<const float4*>(h_s);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
      for (int b = 0; b < kEncB; ++b) {
        const float4 hv = h4[b * (kH / 4) + k];
        acc[b] = fmaf(w[4 * k + 0], hv.x, acc[b]);
        acc[b] = fmaf(w[4 * k + 1], hv.y, acc[b]);
        acc[b] = fmaf(w[4 * k + 2], hv.z, acc[b]);
        acc[b] = fmaf(w[4 * k + 3], hv.w, acc[b]);
      }
    }
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const float w0 = Ws[(4 * k + 0) * kG + j], w1 = Ws[(4 * k + 1) * kG + j];
      const float w2 = Ws[(4 * k + 2) * kG + j], w3 = Ws[(4 * k + 3) * kG + j];
#pragma unroll
      for (int b = 0; b < kEncB; ++b) {
        const float4 hv = h4[b * (kH / 4) + 16 + k];
        acc[b] = fmaf(w0, hv.x, acc[b]);
        acc[b] = fmaf(w1, hv.y, acc[b]);
        acc[b] = fmaf(w2, hv.z, acc[b]);
        acc[b] = fmaf(w3, hv.w, acc[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < kEncB; ++b) g_s[b * kG + j] = acc[b];
    __syncthreads();
#pragma unroll
    for (int p = 0; p < kEncP; ++p) {
#pragma unroll
      for (int q = 0; q < kG / kH; ++q) {   // static slot index: the per-slot arrays stay in registers
        const int bb = q + (kG / kH) * p;
        if ((j >> 7) == q && s < Tz[bb]) {
          const float* g = g_s + bb * kG;
          const float ig = sigmoidf_(g[u]), fg = sigmoidf_(g[kH + u]);
          const float gg = tanhf(g[2 * kH + u]), og = sigmoidf_(g[3 * kH + u]);
          c[p] = fg * c[p] + ig * gg;
          h[p] = og * tanhf(c[p]);
          h_s[bb * kH + u] = h[p];
          const int t = dir ? Tz[bb] - 1 - s : s;
          enc_out[(r0[bb] + t) * (2 * kH) + dir * kH + u] = h[p];
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < kEncP; ++p) {
#pragma unroll
    for (int q = 0; q < kG / kH; ++q) {
      const int bb = q + (kG / kH) * p;
      if ((j >> 7) == q && vid[bb] >= 0) {
        hn[(static_cast<int64_t>(vid[bb]) * 2 + dir) * kH + u] = h[p];
        cn[(static_cast<int64_t>(vid[bb]) * 2 + dir) * kH + u] = c[p];
      }
    }
  }
}

}  // namespace
}  // namespace mucon

// mucon_shead_weights (include/mucon_b200.h): the decoder's parameters, device pointers, fp32, torch layouts
// (Linear weight [out, in]): hid = fs_encoder_hidden_out [H,2H], cn = fs_encoder_cn_out [H,2H], l2 = fs_decoder_attention_l2
// [H,H], att_v [H], emb [C+2,H], comb = fs_decoder_attn_combine [H,3H], wih/whh/bih/bhh = fs_decoder_lstm, t1/t2 =
// fs_decoder_transcript.0/.2 ([H,H], [C+1,H]), n1/n2 = fs_decoder_length.0/.2 ([H/2, H+C+1], [1,H/2]).
namespace mucon {
namespace {

constexpr int kDecThreads = 256;
constexpr int kMaxWords = 128;   // C + 1 <= 128
constexpr int kDecB = 1;         // videos per CTA decoded in lockstep (c2: 1 -> 2.7 ms, 2 -> 3.0 ms, 4 -> 4.3 ms: the decoder is
                                 // bound by the latency of its chain of small phases, not by weight traffic, so more CTAs win)
constexpr int kDecVec = 3 * kH + 3 * kH + (kH + kMaxWords) + 2 * kG + kH;   // floats of per-video vectors

// out[b][r] = act(bias[r] + sum_k W[r, k] * x[b][k]) for r < rows, b < kDecB: a warp takes kDecR rows at a time with the
// lanes across k (coalesced weight reads, every weight used kDecB times).  The kDecR rows' loads are independent, so a
// warp has kDecR x (cols / 32) L2 requests in flight instead of waiting one round trip per row -- with one row at a time
// the decoder spent ~90 % of a step in those round trips.
constexpr int kDecR = 8;
__device__ __forceinline__ void matvecB(float* out, int ldo, const float* __restrict__ W, const float* __restrict__ bias,
                                        const float* x, int ldx, int rows, int cols, bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int rb = warp * kDecR; rb < rows; rb += nw * kDecR) {
    float acc[kDecR][kDecB];
#pragma unroll
    for (int i = 0; i < kDecR; ++i)
#pragma unroll
      for (int b = 0; b < kDecB; ++b) acc[i][b] = 0.f;
    for (int k = lane; k < cols; k += 32) {
      float wv[kDecR], xv[kDecB];
#pragma unroll
      for (int i = 0; i < kDecR; ++i) wv[i] = rb + i < rows ? W[static_cast<int64_t>(rb + i) * cols + k] : 0.f;
#pragma unroll
      for (int b = 0; b < kDecB; ++b) xv[b] = x[b * ldx + k];
#pragma unroll
      for (int i = 0; i < kDecR; ++i)
#pragma unroll
        for (int b = 0; b < kDecB; ++b) acc[i][b] = fmaf(wv[i], xv[b], acc[i][b]);
    }
#pragma unroll
    for (int i = 0; i < kDecR; ++i)
#pragma unroll
      for (int b = 0; b < kDecB; ++b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i][b] += __shfl_xor_sync(0xffffffffu, acc[i][b], o);
      }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kDecR; ++i) {
        if (rb + i >= rows) break;
        const float bv = bias ? bias[rb + i] : 0.f;
#pragma unroll
        for (int b = 0; b < kDecB; ++b) {
          const float v = acc[i][b] + bv;
          out[b * ldo + rb + i] = relu ? fmaxf(v, 0.f) : v;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kDecThreads, 2)
seq_decoder_kernel(const mucon_shead_weights w, const float* __restrict__ enc /*[rows, 2H]*/,
                   const float* __restrict__ enc_ready /*[rows, H]*/, const float* __restrict__ hn,
                   const float* __restrict__ cn, const int64_t* __restrict__ row_off, const int32_t* __restrict__ order,
                   int V, int max_Tz, const int32_t* __restrict__ tf_in, const int32_t* __restrict__ tf_off,
                   int teacher_forcing, int max_steps, int n_words /*C + 1*/, int eos,
                   float* __restrict__ out_logp /*[V, max_steps, n_words]*/, float* __restrict__ out_len /*[V, max_steps]*/,
                   int32_t* __restrict__ out_tokens /*[V, max_steps]*/, int32_t* __restrict__ n_steps /*[V]*/) {
  extern __shared__ __align__(16) float sm[];
  // per video slot b (stride kDecVec): h | c | he | cat (relu(embedding), context) | x (output_attn, logits) | gates x2 | tmp
  float* h = sm;
  float* c = h + kH;
  float* he = c + kH;
  float* cat = he + kH;
  float* x = cat + 3 * kH;
  float* gates = x + kH + kMaxWords;
  float* gates2 = gates + kG;
  float* tmp = gates2 + kG;
  float* red = sm + kDecB * kDecVec;          // [kDecB][8]
  float* scores = red + kDecB * 8;            // [kDecB][max_Tz]
  __shared__ int token_s[kDecB], stop_s[kDecB], steps_s[kDecB], done_s[kDecB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = kDecThreads >> 5;
  int vid[kDecB], Tz[kDecB];
  int64_t r0[kDecB];
  int S = 0;
#pragma unroll
  for (int b = 0; b < kDecB; ++b) {
    const int idx = blockIdx.x * kDecB + b;
    vid[b] = idx < V ? order[idx] : -1;
    r0[b] = vid[b] >= 0 ? row_off[vid[b]] : 0;
    Tz[b] = vid[b] >= 0 ? static_cast<int>(row_off[vid[b] + 1] - r0[b]) : 0;
    const int nst = vid[b] < 0 ? 0 : (teacher_forcing ? min(tf_off[vid[b] + 1] - tf_off[vid[b]], max_steps) : max_steps);
    S = max(S, nst);
    if (tid == 0) {
      steps_s[b] = nst;
      token_s[b] = vid[b] >= 0 ? tf_in[tf_off[vid[b]]] : 0;
      stop_s[b] = 0;
      done_s[b] = 0;
    }
  }
  // decoder initial state from the encoder's final states (models.py:606-622): [h_fwd | h_bwd] -> Linear
#pragma unroll
  for (int b = 0; b < kDecB; ++b)
    for (int k = tid; k < 2 * kH; k += kDecThreads)
      cat[b * kDecVec + k] = vid[b] >= 0 ? hn[static_cast<int64_t>(vid[b]) * 2 * kH + k] : 0.f;
  __syncthreads();
  matvecB(h, kDecVec, w.hid_w, w.hid_b, cat, kDecVec, kH, 2 * kH, false);
  __syncthreads();
#pragma unroll
  for (int b = 0; b < kDecB; ++b)
    for (int k = tid; k < 2 * kH; k += kDecThreads)
      cat[b * kDecVec + k] = vid[b] >= 0 ? cn[static_cast<int64_t>(vid[b]) * 2 * kH + k] : 0.f;
  __syncthreads();
  matvecB(c, kDecVec, w.cn_w, w.cn_b, cat, kDecVec, kH, 2 * kH, false);
  __syncthreads();
  for (int step = 0; step < S; ++step) {
    // which slots take this step (uniform: shared flags written before the last barrier)
    bool act[kDecB];
    bool any = false;
#pragma unroll
    for (int b = 0; b < kDecB; ++b) { act[b] = step < steps_s[b] && !stop_s[b]; any |= act[b]; }
    if (!any) break;
    // embedding -> ReLU (dropout is the identity in eval mode)
#pragma unroll
    for (int b = 0; b < kDecB; ++b) {
      const int token = (teacher_forcing && vid[b] >= 0 && act[b]) ? tf_in[tf_off[vid[b]] + step] : token_s[b];
      for (int k = tid; k < kH; k += kDecThreads) cat[b * kDecVec + k] = fmaxf(w.emb[static_cast<int64_t>(token) * kH + k], 0.f);
    }
    // attention (models.py:731-745): u_t = tanh(enc_ready[t] + l2(h)), a = softmax_t(u_t . V)
    matvecB(he, kDecVec, w.l2_w, w.l2_b, h, kDecVec, kH, kH, false);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < kDecB; ++b) {
      if (!act[b]) continue;
      const float* erv = enc_ready + r0[b] * kH;
      const float* heb = he + b * kDecVec;
      for (int t = warp; t < Tz[b]; t += nw) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < kH / 32; ++q) {
          const int k = lane + 32 * q;
          acc = fmaf(tanhf(erv[static_cast<int64_t>(t) * kH + k] + heb[k]), w.att_v[k], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) scores[b * max_Tz + t] = acc;
      }
    }
    __syncthreads();
    float m[kDecB];
#pragma unroll
    for (int b = 0; b < kDecB; ++b) {
      m[b] = -INFINITY;
      if (act[b]) for (int t = tid; t < Tz[b]; t += kDecThreads) m[b] = fmaxf(m[b], scores[b * max_Tz + t]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m[b] = fmaxf(m[b], __shfl_xor_sync(0xffffffffu, m[b], o));
      if (lane == 0) red[b * 8 + warp] = m[b];
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < kDecB; ++b) {
      m[b] = red[b * 8];
      for (int q = 1; q < nw; ++q) m[b] = fmaxf(m[b], red[b * 8 + q]);
    }
    __syncthreads();
    float ssum[kDecB];
#pragma unroll
    for (int b = 0; b < kDecB; ++b) {
      ssum[b] = 0.f;
      if (act[b]) for (int t = tid; t < Tz[b]; t += kDecThreads) {
        const float e = expf(scores[b * max_Tz + t] - m[b]);
        scores[b * max_Tz + t] = e;
        ssum[b] += e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ssum[b] += __shfl_xor_sync(0xffffffffu, ssum[b], o);
      if (lane == 0) red[b * 8 + warp] = ssum[b];
    }
    __syncthreads();
    // context = sum_t a_t * enc[t]  (thread d owns dimension d: coalesced rows)
#pragma unroll
    for (int b = 0; b < kDecB; ++b) {
      float tot = 0.f;
      for (int q = 0; q < nw; ++q) tot += red[b * 8 + q];
      const float inv = tot > 0.f ? 1.f / tot : 0.f;   // Tz == 0 (a video shorter than the pooling factor): zero context
      float acc = 0.f;
      if (act[b]) {
        const float* encv = enc + r0[b] * (2 * kH);
        for (int t = 0; t < Tz[b]; ++t) acc = fmaf(scores[b * max_Tz + t] * inv, encv[static_cast<int64_t>(t) * (2 * kH) + tid], acc);
      }
      cat[b * kDecVec + kH + tid] = acc;
    }
    __syncthreads();
    matvecB(x, kDecVec, w.comb_w, w.comb_b, cat, kDecVec, kH, 3 * kH, true);      // output_attn = relu(attn_combine(.))
    __syncthreads();
    matvecB(gates, kDecVec, w.wih, w.bih, x, kDecVec, kG, kH, false);
    matvecB(gates2, kDecVec, w.whh, w.bhh, h, kDecVec, kG, kH, false);
    __syncthreads();
    for (int i = tid; i < kDecB * kH; i += kDecThreads) {
      const int b = i >> 7, u = i & (kH - 1);
      if (step < steps_s[b] && !stop_s[b]) {
        const float* g1 = gates + b * kDecVec;
        const float* g2 = gates2 + b * kDecVec;
        const float ig = sigmoidf_(g1[u] + g2[u]), fg = sigmoidf_(g1[kH + u] + g2[kH + u]);
        const float gg = tanhf(g1[2 * kH + u] + g2[2 * kH + u]);
        const float og = sigmoidf_(g1[3 * kH + u] + g2[3 * kH + u]);
        const float cc = fg * c[b * kDecVec + u] + ig * gg;
        c[b * kDecVec + u] = cc;
        h[b * kDecVec + u] = og * tanhf(cc);
      }
    }
    __syncthreads();
    matvecB(tmp, kDecVec, w.t1_w, w.t1_b, h, kDecVec, kH, kH, true);
    __syncthreads();
    matvecB(x + kH, kDecVec, w.t2_w, w.t2_b, tmp, kDecVec, n_words, kH, false);    // transcript logits
    __syncthreads();
    // log_softmax + first-maximum argmax of the transcript logits (thread b for slot b); the length head then reads
    // relu(cat(output_attn, logits))
    if (tid < kDecB) {
      const int b = tid;
      if (vid[b] >= 0 && step < steps_s[b] && !stop_s[b]) {
        const float* xl = x + b * kDecVec + kH;
        float mx = -INFINITY;
        int am = 0;
        for (int q = 0; q < n_words; ++q)
          if (xl[q] > mx) { mx = xl[q]; am = q; }
        float se = 0.f;
        for (int q = 0; q < n_words; ++q) se += expf(xl[q] - mx);
        const float lse = mx + logf(se);
        float* ol = out_logp + (static_cast<int64_t>(vid[b]) * max_steps + step) * n_words;
        for (int q = 0; q < n_words; ++q) ol[q] = xl[q] - lse;
        out_tokens[static_cast<int64_t>(vid[b]) * max_steps + step] = am;
        token_s[b] = am;
        done_s[b] = step + 1;
        if (!teacher_forcing && am == eos) stop_s[b] = 2;   // 2: stop after this step's length has been written
      }
    }
    __syncthreads();
    for (int i = tid; i < kDecB * (kH + n_words); i += kDecThreads) {
      const int b = i / (kH + n_words), k = i - b * (kH + n_words);
      x[b * kDecVec + k] = fmaxf(x[b * kDecVec + k], 0.f);   // (output_attn is >= 0 already)
    }
    __syncthreads();
    matvecB(tmp, kDecVec, w.n1_w, w.n1_b, x, kDecVec, kH / 2, kH + n_words, true);
    __syncthreads();
    matvecB(red, 8, w.n2_w, w.n2_b, tmp, kDecVec, 1, kH / 2, false);
    __syncthreads();
    if (tid < kDecB) {
      const int b = tid;
      if (vid[b] >= 0 && done_s[b] == step + 1) out_len[static_cast<int64_t>(vid[b]) * max_steps + step] = red[b * 8];
      if (stop_s[b] == 2) stop_s[b] = 1;
    }
    __syncthreads();
  }
  if (tid < kDecB && vid[tid] >= 0) n_steps[vid[tid]] = done_s[tid];
}

}  // namespace
}  // namespace mucon

// ---------------------------------------------------------------------------------------------------------------
// fp32 GEMM with bias for the s-head's three projections (W_ih x_t of both directions, enc @ W1):
//   out[m, n] = bias[n] + sum_k A[m, k] * B[k, n]      A [M, K] row-major, B [K, N] row-major, N % 128 == 0, K % 8 == 0
// Exact fp32 on the CUDA cores (the s-head feeds discrete decisions; no TF32): 128 x 128 CTA tile, 256 threads, 8 x 8
// outputs per thread (two LDS.128 of A and two of B per 64 FMAs), k ascending per output.
namespace mucon {
namespace {
constexpr int kSgBM = 128, kSgBN = 128, kSgBK = 8;
__global__ void __launch_bounds__(256)
sgemm_bias_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ bias,
                  float* __restrict__ out, int64_t M, int K, int N) {
  __shared__ __align__(16) float As[2][kSgBK][kSgBM + 4];
  __shared__ __align__(16) float Bs[2][kSgBK][kSgBN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kSgBM;
  const int n0 = blockIdx.y * kSgBN;
  // loader roles: A tile [128 x 8]: thread -> (row tid / 2, k half tid % 2); B tile [8 x 128]: (k tid / 32, col quad tid % 32)
  const int ar = tid >> 1, ak = (tid & 1) * 4;
  const int bk = tid >> 5, bc = (tid & 31) * 4;
  const bool a_ok = m0 + ar < M;
  const float* ap = A + (m0 + ar) * K + ak;
  const float* bp = B + static_cast<int64_t>(bk) * N + n0 + bc;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 av = a_ok ? *reinterpret_cast<const float4*>(ap) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bv = *reinterpret_cast<const float4*>(bp);
  const int nk = K / kSgBK;
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    As[buf][ak + 0][ar] = av.x; As[buf][ak + 1][ar] = av.y; As[buf][ak + 2][ar] = av.z; As[buf][ak + 3][ar] = av.w;
    *reinterpret_cast<float4*>(&Bs[buf][bk][bc]) = bv;
    __syncthreads();
    if (kb + 1 < nk) {   // the next k-block's loads are in flight while this one is multiplied
      av = a_ok ? *reinterpret_cast<const float4*>(ap + (kb + 1) * kSgBK) : make_float4(0.f, 0.f, 0.f, 0.f);
      bv = *reinterpret_cast<const float4*>(bp + static_cast<int64_t>(kb + 1) * kSgBK * N);
    }
#pragma unroll
    for (int k = 0; k < kSgBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    // (double-buffered shared tiles: the stores of block kb + 1 go to the other buffer, one barrier per k-block)
  }
  float bz[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bz[j] = bias ? bias[n0 + tx * 8 + j] : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m < M) {
      float* o = out + m * N + n0 + tx * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[i][0] + bz[0], acc[i][1] + bz[1], acc[i][2] + bz[2], acc[i][3] + bz[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[i][4] + bz[4], acc[i][5] + bz[5], acc[i][6] + bz[6], acc[i][7] + bz[7]);
    }
  }
}
}  // namespace
}  // namespace mucon

extern "C" int mucon_sgemm_bias(const float* A, const float* B, const float* bias, float* out, int64_t M, int K, int N,
                                void* stream) {
  using namespace mucon;
  if (!A || !B || !out || M < 0 || K < 1 || N < 1) return MUCON_EINVAL;
  if (N % kSgBN != 0 || K % kSgBK != 0) return MUCON_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return MUCON_EALIGN;
  if (M == 0) return MUCON_OK;
  const int64_t gx = (M + kSgBM - 1) / kSgBM;
  if (gx > 0x7fffffff) return MUCON_EUNSUPPORTED;
  sgemm_bias_kernel<<<dim3(static_cast<unsigned>(gx), N / kSgBN), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, B, bias, out,
                                                                                                                 M, K, N);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_lstm_encoder(const float* xproj_f, const float* xproj_b, const float* whh_f, const float* whh_b,
                                  const int64_t* row_off, const int32_t* order, int V, int H, float* enc_out, float* hn,
                                  float* cn, void* stream) {
  using namespace mucon;
  if (!xproj_f || !xproj_b || !whh_f || !whh_b || !row_off || !order || !enc_out || !hn || !cn || V < 0)
    return MUCON_EINVAL;
  if (H != kH) return MUCON_EUNSUPPORTED;
  if (V == 0) return MUCON_OK;
  if (V > 65535 * 32) return MUCON_EUNSUPPORTED;
  const int smem = (64 * kG + kEncB * kH + kEncB * kG) * static_cast<int>(sizeof(float));
  MUCON_CUDA_CHECK(cudaFuncSetAttribute(lstm_recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  lstm_recurrent_kernel<<<dim3((V + kEncB - 1) / kEncB, 2), kG, smem, static_cast<cudaStream_t>(stream)>>>(
      xproj_f, xproj_b, whh_f, whh_b, row_off, order, V, enc_out, hn, cn);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_seq_decoder(const mucon_shead_weights* w_h, const float* enc, const float* enc_ready,
                                 const float* hn, const float* cn, const int64_t* row_off, const int32_t* order, int V,
                                 int max_Tz,
                                 const int32_t* tf_in, const int32_t* tf_off, int teacher_forcing, int max_steps,
                                 int n_words, int eos, float* out_logp, float* out_len, int32_t* out_tokens,
                                 int32_t* n_steps, void* stream) {
  using namespace mucon;
  if (!w_h || !enc || !enc_ready || !hn || !cn || !row_off || !order || !tf_in || !tf_off || !out_logp || !out_len ||
      !out_tokens || !n_steps || V < 0 || max_steps < 1 || n_words < 2 || max_Tz < 0)
    return MUCON_EINVAL;
  if (n_words > kMaxWords) return MUCON_EUNSUPPORTED;
  if (V == 0) return MUCON_OK;
  const int mtz = max_Tz > 0 ? max_Tz : 1;
  const int smem = (kDecB * kDecVec + kDecB * 8 + kDecB * mtz + 8) * static_cast<int>(sizeof(float));
  if (smem > 200 * 1024) return MUCON_EUNSUPPORTED;
  MUCON_CUDA_CHECK(cudaFuncSetAttribute(seq_decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  seq_decoder_kernel<<<(V + kDecB - 1) / kDecB, kDecThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      *w_h, enc, enc_ready, hn, cn, row_off, order, V, mtz, tf_in, tf_off, teacher_forcing, max_steps, n_words, eos,
      out_logp, out_len, out_tokens, n_steps);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}
