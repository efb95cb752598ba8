/* oracle.c -- plain-C CPU restatement of the MuCon hot path.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/build.py).  No fast-math: every
 * add/sub/mul below must round exactly as written.
 *
 * Follows (reference paths relative to /root/reference):
 *   orc_block_scores_*  src/core/viterbi/viterbi.py:51 (np.cumsum, sequential) and :68-72
 *   orc_viterbi         src/core/viterbi/viterbi.py:81-138 in the dense form proven equivalent in
 *                       SURVEY.md section 8a (and re-checked by tests/test_oracle_vs_reference.py)
 *   orc_labels          src/core/viterbi/viterbi.py:140-158
 *   orc_poisson_rows    src/core/viterbi/length_model.py:65-80 (given ln m, m, norms per class)
 *   orc_masks           src/mucon/masks.py:19-74 in closed form (affine_grid + bilinear
 *                       grid_sample of a 100-tap template, zero padding)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_INFEASIBLE 1

/* ---- block scores: F[t,c] = F[t-1,c] + logp[t,c] (sequential, in the input dtype);
 *      bs[k,c] = F[fs(k+1)-1,c] - F[fs*k-1,c], bs[0,c] = F[fs-1,c]. */
void orc_block_scores_f32(const float* logp, int64_t T, int C, int fs, float* bs) {
  int64_t K = T / fs;
  float* run = (float*)calloc((size_t)C, sizeof(float));
  float* prev = (float*)calloc((size_t)C, sizeof(float));
  for (int64_t k = 0; k < K; ++k) {
    for (int r = 0; r < fs; ++r) {
      const float* row = logp + ((int64_t)k * fs + r) * C;
      if (k == 0 && r == 0) { for (int c = 0; c < C; ++c) run[c] = row[c]; }
      else { for (int c = 0; c < C; ++c) run[c] = run[c] + row[c]; }
    }
    for (int c = 0; c < C; ++c) {
      bs[k * C + c] = (k == 0) ? run[c] : run[c] - prev[c];
      prev[c] = run[c];
    }
  }
  free(run); free(prev);
}

void orc_block_scores_f64(const double* logp, int64_t T, int C, int fs, double* bs) {
  int64_t K = T / fs;
  double* run = (double*)calloc((size_t)C, sizeof(double));
  double* prev = (double*)calloc((size_t)C, sizeof(double));
  for (int64_t k = 0; k < K; ++k) {
    for (int r = 0; r < fs; ++r) {
      const double* row = logp + ((int64_t)k * fs + r) * C;
      if (k == 0 && r == 0) { for (int c = 0; c < C; ++c) run[c] = row[c]; }
      else { for (int c = 0; c < C; ++c) run[c] = run[c] + row[c]; }
    }
    for (int c = 0; c < C; ++c) {
      bs[k * C + c] = (k == 0) ? run[c] : run[c] - prev[c];
      prev[c] = run[c];
    }
  }
  free(run); free(prev);
}

/* rows[n*J + j-1] = ((l*lnm - m) - lf[l]) - norms, l = j*fs; -inf when l >= max_len.
 * params[n*3 + {0,1,2}] = ln m, m, norms of the label at transcript position n;
 * lf[i] = sum_{k<=i} ln k for i < max_len. */
void orc_poisson_rows(const double* params, int N, const double* lf, int fs, int max_len, double* rows) {
  int J = max_len / fs;
  for (int n = 0; n < N; ++n)
    for (int j = 1; j <= J; ++j) {
      int l = j * fs;
      double v;
      if (l >= max_len) v = -INFINITY;
      else {
        v = (double)l * params[n * 3 + 0];
        v = v - params[n * 3 + 1];
        v = v - lf[l];
        v = v - params[n * 3 + 2];
      }
      rows[n * J + j - 1] = v;
    }
}

/* Dense DP.  bs is [K,C] float (bs_is_f64 = 0) or double.  rows is [N,J].
 * Outputs: *score, seg_blocks[N], bp[K*N] (uint16, 0 = no entry), *jf.  Returns ORC_*. */
int orc_viterbi(const void* bs_, int bs_is_f64, int64_t K, int C, const int32_t* tr, int N,
                const double* rows, int J, int seg0_f32,
                double* score, int32_t* seg_blocks, uint16_t* bp, int32_t* jf_out) {
  if (K < 1 || N < 1 || K > (int64_t)N * J) return ORC_INFEASIBLE;
  const float* bsf = (const float*)bs_;
  const double* bsd = (const double*)bs_;
  size_t sz = (size_t)N * (J + 2);
  double* S = (double*)calloc(sz, sizeof(double));
  double* S2 = (double*)calloc(sz, sizeof(double));
  uint8_t* live = (uint8_t*)calloc(sz, 1);
  uint8_t* live2 = (uint8_t*)calloc(sz, 1);
  memset(bp, 0, sizeof(uint16_t) * (size_t)K * N);
#define AT(n, j) ((size_t)(n) * (J + 2) + (j))
  {
    double b0 = bs_is_f64 ? bsd[tr[0]] : (double)bsf[tr[0]];
    S[AT(0, 1)] = 0.0 + b0; /* exact in either dtype */
    live[AT(0, 1)] = 1;
  }
  for (int64_t k = 1; k < K; ++k) {
    memset(live2, 0, sz);
    for (int n = 0; n < N; ++n) {
      int have = 0; double best = 0.0; int bj = 0;
      for (int j = 1; j <= J; ++j) {
        if (!live[AT(n, j)]) continue;
        double a;
        if (bs_is_f64) a = S[AT(n, j)] + bsd[k * C + tr[n]];
        else if (n == 0 && seg0_f32) a = (double)((float)S[AT(n, j)] + bsf[k * C + tr[n]]);
        else a = S[AT(n, j)] + (double)bsf[k * C + tr[n]];
        if (j < J) { S2[AT(n, j + 1)] = a; live2[AT(n, j + 1)] = 1; }
        if (n + 1 < N) {
          double cand = (a + rows[(size_t)n * J + j - 1]) + 0.0;
          if (!have || best <= cand) { best = cand; bj = j; have = 1; }
        }
      }
      if (have) { S2[AT(n + 1, 1)] = best; live2[AT(n + 1, 1)] = 1; bp[k * N + n + 1] = (uint16_t)bj; }
    }
    { double* t = S; S = S2; S2 = t; uint8_t* u = live; live = live2; live2 = u; }
  }
  memset(seg_blocks, 0, sizeof(int32_t) * (size_t)N);
  if (K < N) {
    *score = -INFINITY; *jf_out = 1;
    for (int n = 0; n < K; ++n) seg_blocks[n] = 1;
  } else {
    int have = 0; double best = 0.0; int bj = 0;
    for (int j = 1; j <= J; ++j) {
      if (!live[AT(N - 1, j)]) continue;
      double cand = (S[AT(N - 1, j)] + rows[(size_t)(N - 1) * J + j - 1]) + 0.0;
      if (!have || best <= cand) { best = cand; bj = j; have = 1; }
    }
    *score = best; *jf_out = bj;
    int n = N - 1; int64_t k0 = K - bj;
    seg_blocks[n] = bj;
    while (n > 0) { int ln = bp[k0 * N + n]; seg_blocks[n - 1] = ln; k0 -= ln; --n; }
  }
#undef AT
  free(S); free(S2); free(live); free(live2);
  return ORC_OK;
}

/* labels[T]: T - fs*K leftover frames first, carrying the last reached segment's label. */
void orc_labels(const int32_t* tr, int N, const int32_t* seg_blocks, int fs, int64_t T, int32_t* labels) {
  int64_t K = T / fs, pos = T - fs * K;
  int last = 0;
  for (int n = 0; n < N; ++n) if (seg_blocks[n] > 0) last = n;
  for (int64_t t = 0; t < pos; ++t) labels[t] = tr[last];
  for (int n = 0; n <= last; ++n)
    for (int64_t e = pos + (int64_t)fs * seg_blocks[n]; pos < e; ++pos) labels[pos] = tr[n];
}

/* One whole video, float32 or float64 log-probs -> labels; the CPU baseline unit. */
int orc_decode_video(const void* logp, int in_is_f64, int64_t T, int C, const int32_t* tr, int N,
                     const double* rows, int J, int fs, int seg0_f32,
                     double* score, int32_t* seg_blocks, int32_t* labels) {
  int64_t K = T / fs;
  if (K < 1) return ORC_INFEASIBLE;
  void* bs = malloc((size_t)K * C * (in_is_f64 ? 8 : 4));
  uint16_t* bp = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)K * N);
  if (in_is_f64) orc_block_scores_f64((const double*)logp, T, C, fs, (double*)bs);
  else orc_block_scores_f32((const float*)logp, T, C, fs, (float*)bs);
  int32_t jf;
  int rc = orc_viterbi(bs, in_is_f64, K, C, tr, N, rows, J, seg0_f32, score, seg_blocks, bp, &jf);
  if (rc == ORC_OK) orc_labels(tr, N, seg_blocks, fs, T, labels);
  free(bs); free(bp);
  return rc;
}

/* ---- masks (float32 arithmetic, torch op order; see oracle/masks.py for the derivation).
 * L is the caller's length vector BEFORE the in-place overlap scaling; on return L_out holds
 * the scaled lengths (the reference mutates its argument, masks.py:61).
 * template: [W] taps; out[M*T]. */
void orc_masks(const float* L, int M, int T, float overlap, const float* tmpl, int W,
               int align_corners, float* L_out, float* out) {
  float cum = 0.f;
  for (int i = 0; i < M; ++i) {
    cum = cum + L[i];
    float pi = cum - L[i];
    float Ls = L[i] * (1.0f + 2 * overlap);
    pi = pi - Ls * (overlap / 2);
    L_out[i] = Ls;
    float s = (float)T / Ls;
    float x = ((pi + Ls / 2) - (float)T / 2) / (-(Ls / 2));
    for (int t = 0; t < T; ++t) {
      float g = align_corners ? (T > 1 ? (2.0f * t) / (T - 1) - 1.0f : 0.f)
                              : (2.0f * t + 1.0f) / T - 1.0f;
      float gx = s * g + x;
      float u = align_corners ? (gx + 1.f) / 2.f * (W - 1) : ((gx + 1.f) * W - 1.f) / 2.f;
      float fl = floorf(u);
      int i0 = (int)fl, i1 = i0 + 1;
      float w1 = u - fl, w0 = 1.f - w1;
      float v = 0.f;
      if (i0 >= 0 && i0 < W) v += tmpl[i0] * w0;
      if (i1 >= 0 && i1 < W) v += tmpl[i1] * w1;
      out[(size_t)i * T + t] = v;
    }
  }
}
