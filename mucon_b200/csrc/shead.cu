// shead.cu -- the sequence-generation ("s") head of the reference at test time (sm_100a).
//
// Replaces, for inference, reference src/mucon/models.py:585-745 (`sequence_generation_forward` and
// `_calculate_attention`): a bidirectional LSTM encoder over the encoded sequence z [Tz, 128], then an attention
// decoder (embedding -> additive attention over the encoder outputs -> attn_combine -> LSTM cell -> transcript head
// and length head) run step by step, teacher-forced or greedy.  The reference runs it one video at a time from Python
// with a device->host `.item()` per decoding step (models.py:721); here a whole batch of variable-length videos is
// two launches and the greedy loop (argmax, EOS test, next input) never leaves the GPU.
//
//   lstm_recurrent_kernel   one CTA per (video, direction): the recurrence h_t = LSTM(Xproj[t] + W_hh h_{t-1}); the
//                           input projections of ALL steps are one conv GEMM launched beforehand (mucon_conv1d).
//                           W_hh (512 x 128 fp32 = 256 KB) lives half in registers (64 per thread, thread j owns gate
//                           row j) and half in shared memory; a step is 128 FMAs per thread and two barriers.
//   seq_decoder_kernel      one CTA per video: all decoding steps; every mat-vec is a warp per output row with the
//                           lanes across the input (coalesced weight reads from L2), fp32 throughout.
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

__global__ void __launch_bounds__(kG, 1)
lstm_recurrent_kernel(const float* __restrict__ xproj_f, const float* __restrict__ xproj_b,
                      const float* __restrict__ whh_f, const float* __restrict__ whh_b,
                      const int64_t* __restrict__ row_off, float* __restrict__ enc_out /*[rows, 2H]*/,
                      float* __restrict__ hn /*[V, 2, H]*/, float* __restrict__ cn /*[V, 2, H]*/) {
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                    // [64][512]: W_hh[j][64 + k] at Ws[k * 512 + j]
  float* h_s = sm + 64 * kG;         // [128]
  float* g_s = h_s + kH;             // [512]
  const int v = blockIdx.x, dir = blockIdx.y, j = threadIdx.x;
  const float* xproj = dir ? xproj_b : xproj_f;
  const float* whh = dir ? whh_b : whh_f;
  const int64_t r0 = row_off[v];
  const int Tz = static_cast<int>(row_off[v + 1] - r0);
  float w[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) w[k] = whh[j * kH + k];
  for (int k = 0; k < 64; ++k) Ws[k * kG + j] = whh[j * kH + 64 + k];
  if (j < kH) h_s[j] = 0.f;
  float c = 0.f, h = 0.f;
  __syncthreads();
  // the input projection of step s + 1 is fetched while step s computes (an L2 round trip per step otherwise)
  float xnext = Tz > 0 ? xproj[(r0 + (dir ? Tz - 1 : 0)) * kG + j] : 0.f;
  for (int s = 0; s < Tz; ++s) {
    const int t = dir ? Tz - 1 - s : s;
    // four independent partial sums (one dependent chain of 128 FMAs would cost 128 x the FMA latency per step)
    float a0 = xnext, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (s + 1 < Tz) xnext = xproj[(r0 + (dir ? t - 1 : t + 1)) * kG + j];
    const float4* h4 = reinterpret_cast<const float4*>(h_s);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 hv = h4[k];
      a0 = fmaf(w[4 * k + 0], hv.x, a0);
      a1 = fmaf(w[4 * k + 1], hv.y, a1);
      a2 = fmaf(w[4 * k + 2], hv.z, a2);
      a3 = fmaf(w[4 * k + 3], hv.w, a3);
    }
#pragma unroll 8
    for (int k = 0; k < 16; ++k) {
      const float4 hv = h4[16 + k];
      a0 = fmaf(Ws[(4 * k + 0) * kG + j], hv.x, a0);
      a1 = fmaf(Ws[(4 * k + 1) * kG + j], hv.y, a1);
      a2 = fmaf(Ws[(4 * k + 2) * kG + j], hv.z, a2);
      a3 = fmaf(Ws[(4 * k + 3) * kG + j], hv.w, a3);
    }
    const float acc = (a0 + a1) + (a2 + a3);
    g_s[j] = acc;
    __syncthreads();
    if (j < kH) {
      const float ig = sigmoidf_(g_s[j]), fg = sigmoidf_(g_s[kH + j]);
      const float gg = tanhf(g_s[2 * kH + j]), og = sigmoidf_(g_s[3 * kH + j]);
      c = fg * c + ig * gg;
      h = og * tanhf(c);
      h_s[j] = h;
      enc_out[(r0 + t) * (2 * kH) + dir * kH + j] = h;
    }
    __syncthreads();
  }
  if (j < kH) {
    hn[(static_cast<int64_t>(v) * 2 + dir) * kH + j] = h;
    cn[(static_cast<int64_t>(v) * 2 + dir) * kH + j] = c;
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

// out[r] = act(b[r] + sum_k W[r, k] * x[k]) for r < rows: a warp per row, lanes across k (coalesced weight reads)
__device__ __forceinline__ void matvec(float* out, const float* __restrict__ W, const float* __restrict__ b,
                                       const float* x, int rows, int cols, bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < rows; r += nw) {
    float acc = 0.f;
    for (int k = lane; k < cols; k += 32) acc = fmaf(W[static_cast<int64_t>(r) * cols + k], x[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      acc += b ? b[r] : 0.f;
      out[r] = relu ? fmaxf(acc, 0.f) : acc;
    }
  }
}

__global__ void __launch_bounds__(kDecThreads)
seq_decoder_kernel(const mucon_shead_weights w, const float* __restrict__ enc /*[rows, 2H]*/,
                   const float* __restrict__ enc_ready /*[rows, H]*/, const float* __restrict__ hn,
                   const float* __restrict__ cn, const int64_t* __restrict__ row_off,
                   const int32_t* __restrict__ tf_in, const int32_t* __restrict__ tf_off, int teacher_forcing,
                   int max_steps, int n_words /*C + 1*/, int eos, float* __restrict__ out_logp /*[V, max_steps, n_words]*/,
                   float* __restrict__ out_len /*[V, max_steps]*/, int32_t* __restrict__ out_tokens /*[V, max_steps]*/,
                   int32_t* __restrict__ n_steps /*[V]*/) {
  extern __shared__ __align__(16) float sm[];
  float* h = sm;                    // [H]
  float* c = h + kH;                // [H]
  float* he = c + kH;               // [H]
  float* cat = he + kH;             // [3H]: relu(embedding) | attention context (2H)
  float* x = cat + 3 * kH;          // [H + kMaxWords]: output_attn | transcript logits (the length head's input)
  float* gates = x + kH + kMaxWords;  // [4H]
  float* gates2 = gates + kG;       // [4H]
  float* tmp = gates2 + kG;         // [H]
  float* red = tmp + kH;            // [32]
  float* scores = red + 32;         // [Tz]
  __shared__ int token_s, stop_s;
  const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = kDecThreads >> 5;
  const int64_t r0 = row_off[v];
  const int Tz = static_cast<int>(row_off[v + 1] - r0);
  const float* encv = enc + r0 * (2 * kH);
  const float* erv = enc_ready + r0 * kH;
  // decoder initial state from the encoder's final states (models.py:606-622): [h_fwd | h_bwd] -> Linear
  for (int k = tid; k < 2 * kH; k += kDecThreads) cat[k] = hn[static_cast<int64_t>(v) * 2 * kH + k];
  __syncthreads();
  matvec(h, w.hid_w, w.hid_b, cat, kH, 2 * kH, false);
  __syncthreads();
  for (int k = tid; k < 2 * kH; k += kDecThreads) cat[k] = cn[static_cast<int64_t>(v) * 2 * kH + k];
  __syncthreads();
  matvec(c, w.cn_w, w.cn_b, cat, kH, 2 * kH, false);
  __syncthreads();
  const int n_tf = tf_off[v + 1] - tf_off[v];
  const int steps = teacher_forcing ? min(n_tf, max_steps) : max_steps;
  if (tid == 0) { token_s = tf_in[tf_off[v]]; stop_s = 0; }
  __syncthreads();
  int done = 0;
  for (int step = 0; step < steps; ++step) {
    const int token = teacher_forcing ? tf_in[tf_off[v] + step] : token_s;
    // embedding -> ReLU (dropout is the identity in eval mode)
    for (int k = tid; k < kH; k += kDecThreads) cat[k] = fmaxf(w.emb[static_cast<int64_t>(token) * kH + k], 0.f);
    // attention (models.py:731-745): u_t = tanh(enc_ready[t] + l2(h)), a = softmax_t(u_t . V)
    matvec(he, w.l2_w, w.l2_b, h, kH, kH, false);
    __syncthreads();
    for (int t = warp; t < Tz; t += nw) {
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < kH / 32; ++q) {
        const int k = lane + 32 * q;
        acc = fmaf(tanhf(erv[static_cast<int64_t>(t) * kH + k] + he[k]), w.att_v[k], acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) scores[t] = acc;
    }
    __syncthreads();
    float m = -INFINITY;
    for (int t = tid; t < Tz; t += kDecThreads) m = fmaxf(m, scores[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int q = 1; q < nw; ++q) m = fmaxf(m, red[q]);
    __syncthreads();
    float ssum = 0.f;
    for (int t = tid; t < Tz; t += kDecThreads) {
      const float e = expf(scores[t] - m);
      scores[t] = e;
      ssum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    if (lane == 0) red[warp] = ssum;
    __syncthreads();
    ssum = 0.f;
    for (int q = 0; q < nw; ++q) ssum += red[q];
    const float inv = ssum > 0.f ? 1.f / ssum : 0.f;   // Tz == 0 (a video shorter than the pooling factor): zero context
    // context = sum_t a_t * enc[t]  (thread d owns dimension d: coalesced rows)
    {
      float acc = 0.f;
      for (int t = 0; t < Tz; ++t) acc = fmaf(scores[t] * inv, encv[static_cast<int64_t>(t) * (2 * kH) + tid], acc);
      cat[kH + tid] = acc;
    }
    __syncthreads();
    matvec(x, w.comb_w, w.comb_b, cat, kH, 3 * kH, true);        // output_attn = relu(attn_combine(.))
    __syncthreads();
    matvec(gates, w.wih, w.bih, x, kG, kH, false);
    matvec(gates2, w.whh, w.bhh, h, kG, kH, false);
    __syncthreads();
    if (tid < kH) {
      const float ig = sigmoidf_(gates[tid] + gates2[tid]), fg = sigmoidf_(gates[kH + tid] + gates2[kH + tid]);
      const float gg = tanhf(gates[2 * kH + tid] + gates2[2 * kH + tid]);
      const float og = sigmoidf_(gates[3 * kH + tid] + gates2[3 * kH + tid]);
      const float cc = fg * c[tid] + ig * gg;
      c[tid] = cc;
      h[tid] = og * tanhf(cc);
    }
    __syncthreads();
    matvec(tmp, w.t1_w, w.t1_b, h, kH, kH, true);
    __syncthreads();
    matvec(x + kH, w.t2_w, w.t2_b, tmp, n_words, kH, false);     // transcript logits
    __syncthreads();
    // log_softmax + first-maximum argmax of the transcript logits; the length head reads relu(cat(output_attn, logits))
    if (tid == 0) {
      float mx = -INFINITY;
      int am = 0;
      for (int q = 0; q < n_words; ++q)
        if (x[kH + q] > mx) { mx = x[kH + q]; am = q; }
      float se = 0.f;
      for (int q = 0; q < n_words; ++q) se += expf(x[kH + q] - mx);
      const float lse = mx + logf(se);
      float* ol = out_logp + (static_cast<int64_t>(v) * max_steps + step) * n_words;
      for (int q = 0; q < n_words; ++q) ol[q] = x[kH + q] - lse;
      out_tokens[static_cast<int64_t>(v) * max_steps + step] = am;
      token_s = am;
      stop_s = (!teacher_forcing && am == eos) ? 1 : 0;
    }
    __syncthreads();
    for (int k = tid; k < kH + n_words; k += kDecThreads) x[k] = fmaxf(x[k], 0.f);   // (output_attn is >= 0 already)
    __syncthreads();
    matvec(tmp, w.n1_w, w.n1_b, x, kH / 2, kH + n_words, true);
    __syncthreads();
    matvec(red, w.n2_w, w.n2_b, tmp, 1, kH / 2, false);
    __syncthreads();
    if (tid == 0) out_len[static_cast<int64_t>(v) * max_steps + step] = red[0];
    done = step + 1;
    if (stop_s) break;   // uniform: read after the barrier
    __syncthreads();
  }
  if (tid == 0) n_steps[v] = done;
}

}  // namespace
}  // namespace mucon

extern "C" int mucon_lstm_encoder(const float* xproj_f, const float* xproj_b, const float* whh_f, const float* whh_b,
                                  const int64_t* row_off, int V, int H, float* enc_out, float* hn, float* cn,
                                  void* stream) {
  using namespace mucon;
  if (!xproj_f || !xproj_b || !whh_f || !whh_b || !row_off || !enc_out || !hn || !cn || V < 0) return MUCON_EINVAL;
  if (H != kH) return MUCON_EUNSUPPORTED;
  if (V == 0) return MUCON_OK;
  if (V > 65535 * 32) return MUCON_EUNSUPPORTED;
  const int smem = (64 * kG + kH + kG) * static_cast<int>(sizeof(float));
  MUCON_CUDA_CHECK(cudaFuncSetAttribute(lstm_recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  lstm_recurrent_kernel<<<dim3(V, 2), kG, smem, static_cast<cudaStream_t>(stream)>>>(xproj_f, xproj_b, whh_f, whh_b,
                                                                                   row_off, enc_out, hn, cn);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}

extern "C" int mucon_seq_decoder(const mucon_shead_weights* w_h, const float* enc, const float* enc_ready,
                                 const float* hn, const float* cn, const int64_t* row_off, int V, int max_Tz,
                                 const int32_t* tf_in, const int32_t* tf_off, int teacher_forcing, int max_steps,
                                 int n_words, int eos, float* out_logp, float* out_len, int32_t* out_tokens,
                                 int32_t* n_steps, void* stream) {
  using namespace mucon;
  if (!w_h || !enc || !enc_ready || !hn || !cn || !row_off || !tf_in || !tf_off || !out_logp || !out_len ||
      !out_tokens || !n_steps || V < 0 || max_steps < 1 || n_words < 2 || max_Tz < 0)
    return MUCON_EINVAL;
  if (n_words > kMaxWords) return MUCON_EUNSUPPORTED;
  if (V == 0) return MUCON_OK;
  const int smem = (3 * kH + 3 * kH + kH + kMaxWords + 2 * kG + kH + 32 + max_Tz + 8) * static_cast<int>(sizeof(float));
  if (smem > 200 * 1024) return MUCON_EUNSUPPORTED;
  MUCON_CUDA_CHECK(cudaFuncSetAttribute(seq_decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  seq_decoder_kernel<<<V, kDecThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      *w_h, enc, enc_ready, hn, cn, row_off, tf_in, tf_off, teacher_forcing, max_steps, n_words, eos, out_logp, out_len,
      out_tokens, n_steps);
  MUCON_CUDA_CHECK(cudaGetLastError());
  return MUCON_OK;
}
