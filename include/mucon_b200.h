/* mucon_b200.h -- C ABI of the B200-native (sm_100a) MuCon hot path.
 *
 * The reference (yassersouri/MuCon) is pure Python and has no FFI; the entry points below are
 * what a ctypes binding placed behind the reference's own call signatures needs (see
 * INTEGRATION.md for the binding and SURVEY.md section 8b for the boundary):
 *
 *   mucon_viterbi_*    replaces core.viterbi.viterbi.Viterbi.decode
 *                      (reference src/core/viterbi/viterbi.py:49-158), called from
 *                      MuConEvaluator.batch_eval_calculation (src/mucon/evaluators.py:178-180)
 *   mucon_masks_*      replaces mucon.masks.create_masks (src/mucon/masks.py:19-74), called from
 *                      MuCon.mucon_loss (src/mucon/models.py:430-441)
 *   mucon_backbone_*   replaces WaveNetBlock.forward (src/core/modules/temporal.py:128-147) +
 *                      MuCon.temporal_modeling_forward tail (src/mucon/models.py:759-768) +
 *                      frame_classifier_forward / log_softmax (src/mucon/models.py:567-582, :368)
 *
 * Conventions: plain pointers and sizes only.  Every pointer is a DEVICE pointer unless its name
 * ends in _h.  Every call is asynchronous on `stream` (a cudaStream_t passed as void*), returns
 * 0 or a negative MUCON_E* code, never throws.  Global state is limited to read-only caches (the device's SM
 * count, the three constant mask templates, the last CUDA error string).  The caller owns all buffers.  There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MUCON_B200_H_
#define MUCON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUCON_ABI_VERSION 2

/* call-level errors (return values) */
#define MUCON_OK 0
#define MUCON_EINVAL (-1)      /* bad argument (null pointer, non-positive size, ...) */
#define MUCON_EUNSUPPORTED (-2) /* shape outside what the kernels cover (J > 128, N > 65, ...) */
#define MUCON_ECUDA (-3)       /* a CUDA runtime call failed; see mucon_last_cuda_error() */
#define MUCON_EALIGN (-4)      /* pointer not aligned as documented */
#define MUCON_ESHAPE (-5)      /* input larger than the session / workspace was created for */

/* per-unit status written by mucon_viterbi_decode (int32 each) */
#define MUCON_UNIT_OK 0
#define MUCON_UNIT_INFEASIBLE 1 /* T < fs or K > N*J: the reference raises (SURVEY.md V-edge) */
#define MUCON_UNIT_SHORT 2      /* K < N: score -inf, partial path, same as the reference */
#define MUCON_UNIT_NONFINITE 3  /* best score is NaN or +inf */

int mucon_abi_version(void);
/* Caps the grid of the persistent kernels launched by the calling host thread at n CTAs (0 = no cap); returns the
 * previous cap.  Two capped kernels on two streams then run side by side on disjoint SMs. */
int mucon_set_sm_limit(int n);
const char* mucon_strerror(int code);
const char* mucon_last_cuda_error(void);
/* Number of SMs / device name of the current device; 0 / "" without a device. */
int mucon_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Viterbi, step 1: block scores.
 * For every video v (rows vid_off[v] .. vid_off[v+1] of logp, row-major [rows, C]):
 *   F[t,c] = F[t-1,c] + logp[t,c]   sequentially in the input dtype   (viterbi.py:51)
 *   bs[k,c] = F[fs(k+1)-1,c] - F[fs*k-1,c],  bs[0,c] = F[fs-1,c]      (viterbi.py:68-72)
 * written to rows blk_off[v] .. blk_off[v+1] of bs ([rows, C], same dtype as logp), with
 * blk_off[v+1]-blk_off[v] == (vid_off[v+1]-vid_off[v]) / fs.
 * order: optional [V] permutation (launch order, longest first); NULL = identity.
 * logp must be 16-byte aligned.
 */
int mucon_viterbi_blockscores(const void* logp, int in_is_f64,
                              const int64_t* vid_off, const int64_t* blk_off, const int32_t* order,
                              int V, int C, int fs, void* bs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Viterbi, step 2: dynamic program + traceback + label writer, one CTA per unit
 * (unit = one video x one candidate transcript).
 */
typedef struct mucon_viterbi_batch {
  int32_t U;          /* units */
  int32_t C;          /* classes (row width of bs) */
  int32_t fs;         /* frame_sampling (viterbi.py:34; the evaluator uses 30) */
  int32_t max_len;    /* length_model.max_length() (length_model.py:82; 2000); max_len/fs <= 128 */
  int32_t bs_is_f64;  /* dtype of bs */
  int32_t seg0_f32;   /* 1: segment 0 accumulates in float32 (NumPy>=2 promotion) */
  int32_t max_N;      /* max transcript length over units (see mucon_viterbi_pack_h) */
  int32_t max_K;      /* max number of blocks over units (sizes the shared back-pointer stage) */
  int32_t n_cta;      /* number of unit bins, from mucon_viterbi_pack_h */
  int32_t wpc;        /* warps per CTA (4, 8 or 16), from mucon_viterbi_pack_h */
  int32_t lanes;      /* lanes per segment (4, 8 or 32), from mucon_viterbi_pack_h */
  int32_t n_peers;    /* 0, or the number of valid entries of peer_delta (multi-GPU result exchange, see below) */
  const void* bs;            /* [sum K, C] block scores */
  const int64_t* vid_off;    /* [V+1] frame offsets (T_v = difference) */
  const int64_t* blk_off;    /* [V+1] block offsets into bs */
  const int32_t* unit_vid;   /* [U] video of each unit */
  const int32_t* tr;         /* [sum N] transcripts, concatenated */
  const int32_t* tr_off;     /* [U+1] */
  const double* len_rows;    /* [sum N, J] length scores of j=1..J blocks (J = max_len/fs), or NULL */
  const double* len_params;  /* [sum N, 3] (ln m, m, norms) per transcript position, or NULL */
  const double* logfact;     /* [J+1] sum_{i<=j*fs} ln i at index j; used with len_params */
  const int64_t* lab_off;    /* [U] offset of the unit's labels in `labels`, <0 = do not write */
  const int64_t* bp_off;     /* [U] offset (bytes) of the unit's [K,N] uint8 back-pointer table */
  const int32_t* warp_unit;  /* [n_cta*wpc] unit of every warp of every bin (-1 = unused) */
  double* score;             /* [U] */
  int32_t* labels;           /* frame labels, T_u each */
  int32_t* seg_blocks;       /* [sum N] segment lengths in blocks (0 = segment not reached) */
  uint8_t* bp;               /* back-pointers: winning predecessor length, 0 = no entry */
  int32_t* final_j;          /* [U] length (blocks) of the last segment */
  int32_t* status;           /* [U] MUCON_UNIT_* */
  /* Multi-GPU result exchange fused into the kernels' epilogue (replaces the NCCL all_gather of the per-video
   * scores and segment lengths): `score` and `seg_blocks` live in one payload buffer; every store to
   * them is repeated at  (char*)address + peer_delta[p]  for p < n_peers -- this rank's slot in the receive buffer
   * of rank p, mapped into this process with mucon_peer_open (peer-to-peer stores over NVLink). */
  int64_t peer_delta[8];
} mucon_viterbi_batch;

/* Host helper: packs units into bins of wpc warps (one CTA each; wpc = 4, 8 or 16, the
 * smallest that holds the largest unit).  `lanes` = lanes per transcript segment: 32 gives every
 * segment its own warp (lowest latency per DP step; needs max_N <= 17), 0/4/8 lets 8 (J <= 32)
 * or 4 segments share a warp (fewest instructions; N <= 129 / 65).  The choice made is returned
 * in *lanes_out and must be passed on in mucon_viterbi_batch.  A unit needs
 * max(1, ceil((N-1)/(32/lanes))) consecutive warps.  Units are taken in order_h (or 0..U-1) --
 * pass them longest first.  warp_unit_h needs room for U*16 entries; on return the first
 * *n_cta_out * *wpc_out are valid. */
int mucon_viterbi_pack_h(const int32_t* N_h, const int32_t* order_h, int U, int max_N, int fs,
                         int max_len, int lanes, int32_t* warp_unit_h, int32_t* n_cta_out,
                         int32_t* wpc_out, int32_t* lanes_out);

/* ---- the sequence-generation ("s") head at test time (src/mucon/models.py:585-745) -------------------------------
 * mucon_lstm_encoder: the recurrence of the bidirectional LSTM encoder for a packed batch.  xproj_f / xproj_b
 * [rows, 4H]: W_ih x_t + b_ih + b_hh of every step for the forward / reverse direction (one mucon_conv1d launch each),
 * whh_f / whh_b [4H, H] (torch gate order i, f, g, o), row_off [V+1], order [V] (a permutation of the videos, longest
 * first: a CTA runs four consecutive entries in lockstep) -> enc_out [rows, 2H] (= fs_encoder_lstm_out), hn / cn
 * [V, 2, H] (final states).  H = 128.
 * mucon_seq_decoder: the attention decoder, every decoding step of every video in one launch.  enc = enc_out,
 * enc_ready [rows, H] = enc @ fs_decoder_attention_W1; order [V]: a permutation of the videos (a CTA decodes four
 * consecutive entries in lockstep: sort by the number of steps); tf_in / tf_off [V+1]: per video SOS + transcript
 * (teacher_forcing != 0: step s is fed tf_in[s]; == 0: greedy, the first-maximum argmax of a step feeds the next and
 * the video stops at `eos` or after max_steps -- the reference's per-step .item(), models.py:721, stays on the device).
 * out_logp [V, max_steps, n_words] log-softmaxed transcript logits, out_len [V, max_steps] length logits,
 * out_tokens [V, max_steps] argmax tokens, n_steps [V] steps taken. */
typedef struct mucon_shead_weights {
  const float *hid_w, *hid_b, *cn_w, *cn_b, *l2_w, *l2_b, *att_v, *emb, *comb_w, *comb_b, *wih, *whh, *bih, *bhh,
      *t1_w, *t1_b, *t2_w, *t2_b, *n1_w, *n1_b, *n2_w, *n2_b;
} mucon_shead_weights;
/* out[m, n] = bias[n] + sum_k A[m, k] * B[k, n] in exact fp32 (CUDA cores): the s-head's input / attention projections.
 * A [M, K], B [K, N] row-major, N % 128 == 0, K % 8 == 0, bias may be NULL. */
int mucon_sgemm_bias(const float* A, const float* B, const float* bias, float* out, int64_t M, int K, int N, void* stream);
int mucon_lstm_encoder(const float* xproj_f, const float* xproj_b, const float* whh_f, const float* whh_b,
                       const int64_t* row_off, const int32_t* order, int V, int H, float* enc_out, float* hn, float* cn,
                       void* stream);
int mucon_seq_decoder(const mucon_shead_weights* w_h, const float* enc, const float* enc_ready, const float* hn,
                      const float* cn, const int64_t* row_off, const int32_t* order, int V, int max_Tz, const int32_t* tf_in,
                      const int32_t* tf_off, int teacher_forcing, int max_steps, int n_words, int eos, float* out_logp,
                      float* out_len, int32_t* out_tokens, int32_t* n_steps, void* stream);

/* Evaluator glue on the device (src/mucon/evaluators.py:155-165, core/viterbi/length_model.py:54-63): from the
 * s-head's relative lengths rel [sum N] (float32), one transcript per video (tr, tr_off [V+1]) and the frame offsets
 * vid_off [V+1], the Poisson parameters (ln m, m, norms) of every transcript position [sum N, 3] -- exactly what
 * mucon_viterbi_batch.len_params takes.  logtail[i] = sum_{k=2..i} ln k (host-built, logtail_n entries, must cover the
 * longest video).  ln() is evaluated on the device: scores can differ from host-built parameters in the last bits. */
int mucon_class_mean_params(const float* rel, const int32_t* tr, const int32_t* tr_off, const int64_t* vid_off, int V,
                            const double* logtail, int logtail_n, double* len_params_out, void* stream);

/* Single-video session -- the reference's call pattern (src/mucon/evaluators.py:147-180: one Viterbi.decode per
 * video, core/viterbi/viterbi.py:49-158) as ONE call from host arrays to host arrays: a pinned staging buffer
 * carries [metadata | log-probabilities] to the device in one copy, the fused alignment kernel runs, one copy
 * brings [score | final_j | status | segment lengths | labels] back; the call returns after the stream has been
 * synchronised.  A session holds buffers for up to max_T frames, C classes, max_N segments of elem_bytes (4 / 8)
 * log-probabilities; ESHAPE asks the caller for a larger session, EUNSUPPORTED for the general entry points.
 * len_params_h: [N,3] (ln m, m, norms) of the transcript's classes in transcript order.  Not thread-safe. */
typedef struct mucon_single mucon_single;
int mucon_single_create(int max_T, int C, int max_N, int elem_bytes, mucon_single** out);
int mucon_single_destroy(mucon_single* s);
int mucon_single_decode_h(mucon_single* s, const void* logp_h, int is_f64, int T, const int32_t* tr_h, int N,
                          const double* len_params_h, int fs, int max_len, int seg0_f32, double* score_h,
                          int32_t* labels_h, int32_t* seg_blocks_h, int32_t* status_h, int32_t* final_j_h, void* stream);

/* Receive buffers of the result exchange: device memory allocated with cudaMalloc (one allocation = one CUDA IPC
 * handle), exported as a 64-byte cudaIpcMemHandle_t, opened by the other ranks of the node (peer access is enabled
 * lazily by the driver), closed / freed at the end.  The handles travel through torch.distributed (or any channel). */
int mucon_peer_alloc(size_t bytes, void** dev_ptr_out, unsigned char handle_out[64]);
int mucon_peer_open(const unsigned char handle[64], void** dev_ptr_out);
int mucon_peer_close(void* dev_ptr);
int mucon_peer_free(void* dev_ptr);

int mucon_viterbi_decode(const mucon_viterbi_batch* batch_h, void* stream);

/* The same dynamic program for shapes the register-resident kernels reject with
 * MUCON_EUNSUPPORTED: J = max_len/fs > 128 (the reference class's own defaults, frame_sampling = 1
 * and max_length = 2000: viterbi.py:34, length_model.py:82) or transcripts beyond 65 segments.
 * One CTA per unit; hypothesis scores live in the caller's workspace: ws_off[u] (in doubles) is
 * the start of 2 * N_u * J doubles for unit u.  bp_is_u16 != 0: batch.bp is a uint16 table
 * (needed when J > 255), bp_off in elements either way.  warp_unit / n_cta / wpc / lanes of the
 * batch are ignored.  Same outputs, bit for bit, as mucon_viterbi_decode where both apply. */
int mucon_viterbi_decode_generic(const mucon_viterbi_batch* batch_h, double* ws, const int64_t* ws_off,
                                 int bp_is_u16, void* stream);

/* Lane-per-segment dynamic program (J <= 66, N <= 33): every lane of a warp owns one transcript
 * segment with its J hypothesis scores in registers, a warp carries several units side by side.
 * mucon_viterbi_pack_lanes_h assigns lanes on the host: a unit takes max(1, N-1) consecutive lanes
 * of one warp (units in order_h, longest first); lane_unit_h needs room for U*32 entries, on
 * return the first *n_warps_out * 32 are valid (-1 = unused lane).
 * mucon_viterbi_decode_lanes: same inputs/outputs as mucon_viterbi_decode (warp_unit / n_cta /
 * wpc / lanes ignored), one 32-thread CTA per warp of the packing.  reserved: pass NULL (must be NULL;
 * a scan-concurrent variant that polled per-video progress counters was never shipped). */
int mucon_viterbi_pack_lanes_h(const int32_t* N_h, const int32_t* order_h, int U, int32_t* lane_unit_h,
                               int32_t* n_warps_out);
int mucon_viterbi_decode_lanes(const mucon_viterbi_batch* batch_h, const int32_t* lane_unit, int n_warps,
                               const int32_t* reserved, void* stream);

/* One-launch alignment: block-score scan and DP of a unit fused in one CTA (scan warps feed the
 * DP warps through shared memory; block scores do not travel through HBM).  Same inputs and
 * outputs as mucon_viterbi_blockscores + mucon_viterbi_decode; warp_unit / n_cta / wpc of the
 * batch are ignored; batch.U CTAs are launched for the units order[0..U) (so a subset can be
 * aligned); batch.lanes == 32 gives every transcript segment its own warp (lower latency per DP
 * step, for long videos; needs max_N <= 15), anything else shares a warp between 4 or 8 segments.  Meant for one transcript per video (a unit re-scans its video).
 * logp: [sum T, C] log-probabilities, 16-byte aligned, dtype given by in_is_f64 (must match
 * batch.bs_is_f64).  order: optional [U] launch order (longest first).  write_bs != 0 also stores
 * the block scores to batch.bs.  Returns MUCON_EUNSUPPORTED when the shape does not fit
 * (C*sizeof(dtype) % 16 != 0, C > 128, more than 6 DP warps per unit, ...): call the two-step
 * path instead. */
int mucon_viterbi_align_fused(const mucon_viterbi_batch* batch_h, const void* logp, int in_is_f64,
                              const int32_t* order, int write_bs, void* stream);

/* mucon_viterbi_align_fused with the long tail split off: the first n_wide units of `order` (the
 * longest videos; their DP is the critical path of the launch) run with a warp per segment, the
 * others with the shared-warp kernel, CONCURRENTLY in the same stream: the second kernel is a
 * programmatic dependent launch that starts as soon as the CTAs of the first are resident (the two
 * work on disjoint units).  n_wide == 0 is mucon_viterbi_align_fused.  If the wide shape is not
 * covered everything runs in the main launch. */
int mucon_viterbi_align_fused_tail(const mucon_viterbi_batch* batch_h, const void* logp, int in_is_f64,
                                   const int32_t* order, int n_wide, int write_bs, void* stream);
/* The same alignment fed from the backbone's POOLED resolution: logp_z holds log-probabilities [sum Tz, C]
 * (z_off[V+1] row offsets, device) and frame t of video v reads row min(floor(t * (float)Tz_v / T_v), Tz_v - 1)
 * -- the nearest-neighbour index of F.interpolate (src/mucon/models.py:574-576; the 1x1 classifier and the
 * log-softmax commute with it).  The scan performs the same sequence of float additions as over the expanded
 * [sum T, C] array (np.cumsum, viterbi.py:51), so every output is bit-identical to
 * mucon_logsoftmax_expand + mucon_viterbi_align_fused_tail, without the 4*T*C bytes per video being written
 * or read. */
int mucon_viterbi_align_fused_pooled(const mucon_viterbi_batch* batch_h, const void* logp_z, int in_is_f64,
                                     const int64_t* z_off, const int32_t* order, int n_wide, int write_bs,
                                     void* stream);

/* Arg-max over the candidates of each video: best[v] = unit index with the highest score among
 * units cand_off[v] .. cand_off[v+1] (lowest index wins ties; units with status INFEASIBLE are
 * skipped; -1 if none). */
int mucon_viterbi_select(const double* score, const int32_t* status, const int32_t* cand_off,
                         int V, int32_t* best, void* stream);

/* The same arg-max with the reference's order between candidates whose scores tie EXACTLY: its hypothesis dict is
 * insertion-ordered and finalize_decoding (viterbi.py:125-138) takes `score >= best`, so the final hypothesis latest
 * in the dict wins.  That order is structural: larger tie_rank[2u] (place of the transcript's first label in the
 * grammar's successor iteration), then larger final_j[u] + transcript length, then larger tie_rank[2u + 1] (dense
 * rank of the rest of the transcript's successor places, a proper prefix first); mucon_b200/grammar.py:tie_ranks
 * builds both from the grammar's own successor sets.  Remaining ties (duplicate transcripts): lowest index. */
int mucon_viterbi_select_ranked(const double* score, const int32_t* status, const int32_t* cand_off, int V,
                                const int32_t* final_j, const int32_t* tr_off, const int32_t* tie_rank,
                                int32_t* best, void* stream);

/* Labels for selected units from their seg_blocks: sel[i] is a unit index (or <0 to skip) whose
 * labels are written at labels + out_off[i]. */
int mucon_viterbi_labels(const int32_t* sel, int n_sel, const int64_t* out_off,
                         const int64_t* vid_off, const int32_t* unit_vid, const int32_t* tr,
                         const int32_t* tr_off, const int32_t* seg_blocks, int fs,
                         int32_t* labels, void* stream);

/* Host helper: Poisson parameters (ln m, m, norms) for `n` means with libm's log
 * (length_model.py:54-63).  NumPy callers should build them with numpy instead so that ln()
 * is the very function the reference used. */
int mucon_poisson_params_h(const double* means_h, int n, double* out_h);
/* Host helper: logfact[j] = sum_{i=1..j*fs} ln i for j = 0..J (sequential), J = max_len/fs. */
int mucon_logfact_h(int fs, int max_len, double* out_h);

/* ---------------------------------------------------------------------------------------------
 * Masks (src/mucon/masks.py:19-74).  For video v with M_v = n_off[v+1]-n_off[v] lengths and
 * target size T_v: out rows n_off[v].. are [M_v, T_v] float32 written at out + out_off[v].
 * L is NOT modified; the scaled lengths L*(1+2*overlap) (the reference's in-place side effect,
 * masks.py:61) are written to L_scaled when non-NULL.
 * template_id: 0 box, 1 gaussian(std=20), 2 trapezoid.  align_corners: 0 (torch >= 1.3 default)
 * or 1 (torch 1.1, the reference's pinned docker).
 * row_vid: [n_rows] video of every mask row (one CTA per row looks its video up in one load), or NULL
 * (the kernel then searches n_off).
 */
int mucon_masks_fwd(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* out_off,
                    const int32_t* row_vid, int V, int n_rows /* n_off[V] */, int max_T /* max over T[] */,
                    float overlap, int template_id, int align_corners, float* L_scaled, float* out, void* stream);
/* mucon_masks_fwd in two launches: a thread per row computes the row's geometry (prefix sum of the lengths, window,
 * certified constant ranges) into ws (64 bytes per row, 16-byte aligned), then 64-thread groups write the rows from
 * those records without any dependent metadata loads or barriers.  Same results as mucon_masks_fwd. */
int mucon_masks_fwd_ws(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* out_off,
                       const int32_t* row_vid, int V, int n_rows, int max_T, float overlap, int template_id,
                       int align_corners, float* L_scaled, float* out, void* ws, void* stream);
/* grad_L[i] = d(sum(grad_out * masks))/dL[i], through pi (cumsum) and the scale.
 * grad_out has the layout of `out`; ws is scratch of 2*n_rows floats. */
int mucon_masks_bwd(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* out_off,
                    const int32_t* row_vid, int V, int n_rows, float overlap, int template_id, int align_corners,
                    const float* grad_out, float* ws, float* grad_L, void* stream);
/* mucon_flint_fwd with a warp per (mask row, eighth of its window) instead of a CTA per row: no block barriers, the
 * long windows no longer set the tail.  ws: mucon_flint_fwd_ws_words(n_rows, C) 4-byte words (16-byte aligned):
 * [partial sums | reserved | one 64-byte geometry record per row], overwritten.  Three launches: the rows' geometry (a
 * thread per row), the partial sums (a warp per item), and their addition in a fixed order (a thread per (row, four
 * classes)), so E does not depend on scheduling. */
int64_t mucon_flint_fwd_ws_words(int n_rows, int C);
int mucon_flint_fwd_ws(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* seg_off,
                       const int32_t* row_vid, int V, int n_rows, int C, float overlap, int template_id,
                       int align_corners, const float* seg, float* ws, float* E, void* stream);
/* Fused "flint" evidence of the mutual-consistency loss (models.py:456-468):
 *   E[r, c] = sum_t mask_r[t] * seg[t, c]        r = mask row (video v, segment i), c < C <= 128
 * without materialising the masks; seg is the packed [sum T, C] frame-logit tensor, seg_off[v] the
 * first frame of video v.  A row only reads the frames of its own window.
 * Backward: grad_seg[t, :] = sum_r mask_r[t] * grad_E[r, :] (skipped when grad_seg is NULL; chunks:
 * device array of {int32 video, int32 t0}, one entry per 512 frames of a video) and grad_L through
 * d mask_r[t] = grad_E[r, :] . seg[t, :] (ws: 2*n_rows floats; max_rows = most rows of a video, <= 64;
 * C a multiple of 4). */
int mucon_flint_fwd(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* seg_off,
                    const int32_t* row_vid, int V, int n_rows, int C, float overlap, int template_id,
                    int align_corners, const float* seg, float* E, void* stream);
int mucon_flint_bwd(const float* L, const int32_t* n_off, const int32_t* T, const int64_t* seg_off,
                    const int32_t* row_vid, int V, int n_rows, int max_rows, int C, float overlap, int template_id,
                    int align_corners, const float* seg, const float* grad_E, const void* chunks, int n_chunks,
                    float* grad_seg, float* ws, float* grad_L, void* stream);
/* Host helper: the 100 template taps the kernels use (masks.py:34-54). */
int mucon_mask_template_h(int template_id, float* out100_h);

/* ---------------------------------------------------------------------------------------------
 * Backbone forward (inference).  Activations are time-major with channels contiguous: a video
 * is a [T, C] block of rows and videos are concatenated; row_off[V+1] are the row offsets at the
 * resolution of the call.  The host mirror (mucon_b200/temporal.py, model.py) strings these ops
 * together exactly like WaveNetBlock.forward (src/core/modules/temporal.py:128-147),
 * temporal_modeling_forward (src/mucon/models.py:746-773), frame_classifier_forward (:567-582)
 * and predict's log_softmax (:368).
 */
/* out[m, n] = act(sum_k A[m,k]*W[n,k] + bias[n]), A [M,K] fp32 row-major, W [N,K] fp32 (a Conv1d
 * weight [out, in, 1]), N must be 128, K a multiple of 32.  tcgen05.mma.kind::tf32 (operands are
 * read as TF32: 10-bit mantissa, fp32 accumulate).  Replaces first_conv + ReLU (temporal.py:133). */
int mucon_gemm_tf32_bias_act(const float* A, int64_t M, int K, const float* W, int N, const float* bias,
                             float* out, int relu, void* stream);
/* The same projection with the result stored as bf16 (fp16 != 0: IEEE half, saturating) ([M, 128] row-major,
 * 32-byte aligned): the input of the 16-bit layer kernel below.  Operands are still the fp32 features read as TF32. */
int mucon_gemm_tf32_bias_act_bf16(const float* A, int64_t M, int K, const float* W, int N, const float* bias,
                                  void* out16, int relu, int fp16, void* stream);
/* The same convolution for 128 -> 128 channels on tcgen05 (TF32 operands, fp32 accumulate): every
 * tap is four k-blocks of one TMEM accumulator, A tiles are TMA loads of the time-major activations
 * at a shifted row, rows outside the video are zeroed in shared memory (Conv1d zero padding).
 *   out[t,:] = relu_final( relu_mid( sum_tap in[t + (tap - taps/2)*dilation, :] . W[tap]^T + bias ) + residual[t,:] )
 * W_kco: Conv1d weight [Cout, Cin, k] permuted to [k][Cout][Cin].  tiles: device array of
 * {int64 row0; int32 t0; int32 T} (16 bytes each): one entry per 128-row tile of a video (row0 =
 * first row of the video at this resolution, t0 = tile start within it, T = its length). */
int mucon_conv_gemm_tf32(const float* in, float* out, const float* W_kco, const float* bias,
                         const float* residual, const void* tiles, int num_tiles, int64_t rows, int taps,
                         int dilation, int relu_mid, int relu_final, void* stream);
/* The same kernel with an explicit list of tap row-shifts (<= 6, one of them 0):
 *   out[t,:] = relu_final( relu_mid( sum_i in[t + shifts_h[i], :] . W[i]^T + bias ) + residual[t,:] )
 * W_kco: [n_shifts][Cout][Cin].  Used for the MS-TCN++ first stage (temporal.py:150-204), whose two
 * dilated convolutions and 1x1 fusion conv are linear up to the ReLU and fold into one 5-tap conv. */
int mucon_conv_gemm_tf32_shifts(const float* in, float* out, const float* W_kco, const float* bias,
                                const float* residual, const void* tiles, int num_tiles, int64_t rows,
                                const int32_t* shifts_h, int n_shifts, int relu_mid, int relu_final, void* stream);
/* max_pool1d(2) (mode 0) or avg_pool1d(2) * 2 = the sum of the pair (mode 1; pooling_type != "max",
 * temporal.py:139-142) over time-major rows, floor halving per video.  The relu / relu_mid / relu_final / relu_in /
 * relu_out flags of the backbone entry points are activation modes: 0 none, 1 ReLU, 2 leaky ReLU (slope 0.01,
 * model.ft.leaky_relu, temporal.py:35-41). */
int mucon_pool2(const float* in, float* out, const int64_t* off_in, const int64_t* off_out, int V, int max_T_out,
                int C, int mode, void* stream);
/* conv_gemm with the full epilogue (training step):
 *   out = gate( relu_final( relu_mid( sum_i in[t + shifts_h[i], :] . W[i]^T (+ bias) ) (* mul) (+ residual) ) )
 * bias, residual, mul, gate may be NULL.  mul [rows,128]: the dropout mask/scale of WaveNetLayer.drop
 * (temporal.py:51); gate [rows,128]: out is zeroed where gate <= 0 (the ReLU derivative of the backward pass).
 * The data gradient of a convolution is this call with the taps' [Cout][Cin] slices transposed and the shifts negated. */
int mucon_conv_gemm_tf32_ex(const float* in, float* out, const float* W_kco, const float* bias, const float* residual,
                            const float* mul, const float* gate, const void* tiles, int num_tiles, int64_t rows,
                            const int32_t* shifts_h, int n_shifts, int relu_mid, int relu_final, void* stream);
/* Weight and bias gradients of a 128-output-channel convolution over time-major rows on tcgen05 (TF32 operands read
 * MN-major from the activations as they lie in memory, fp32 accumulation in TMEM), what autograd computes for
 * temporal.py:43-53,133,145 under trainers.py:125-131:
 *   dW[out_off_h[j] + co*ldo + ci] += sum_t dY[t, co] * X[t + shifts_h[j], xcol_h[j] + ci]   (both frames in the video)
 *   dbias[co] += sum_t dY[t, co]                                                              (dbias may be NULL)
 * dY [rows,128], X [rows,ldx] (ldx a multiple of 32, >= 128), n_jobs <= 16 (three taps of a dilated conv, one tap of a
 * 1x1 conv, or the sixteen 128-column blocks of the 2048-d features for first_conv).  dW / dbias are ACCUMULATED
 * (red.global.add): zero them first.  tiles: as mucon_conv_gemm_tf32.  With dbias one of the first four jobs must
 * have shift 0. */
int mucon_wgrad_tf32(const float* dY, const float* X, int ldx, const void* tiles, int num_tiles, int64_t rows,
                     const int32_t* shifts_h, const int32_t* xcol_h, const int64_t* out_off_h, int n_jobs, int ldo,
                     float* dW, float* dbias, void* stream);
/* Backward of mucon_groupnorm_relu (models.py:759-768 under autograd), 128 channels: dx, and dgamma / dbeta
 * ACCUMULATED into zeroed buffers.  x is the GroupNorm input, dy the gradient of relu(gn(x)). */
int mucon_groupnorm_relu_bwd(const float* x, const float* dy, const float* gamma, const float* beta,
                             const int64_t* row_off, int V, int C, int groups, float eps, int relu, float* dx,
                             float* dgamma, float* dbeta, void* stream);
/* Backward of mucon_expand_rows (F.interpolate(mode="nearest"), models.py:574-577): grad_table[iz,:] = sum of
 * grad_out[t,:] over the frames whose source row is iz. */
int mucon_expand_rows_bwd(const float* grad_out, const int64_t* off_z, const int64_t* off_t, int V, int max_Tz, int C,
                          float* grad_table, void* stream);
/* max_pool1d(2) backward (temporal.py:137-139): dx[2t or 2t+1] = dy[t] at the first maximum, 0 elsewhere. */
int mucon_maxpool2_bwd(const float* x, const float* dy, const int64_t* off_in, const int64_t* off_out, int V,
                       int max_T_out, int C, float* dx, void* stream);
/* One whole WaveNet layer (temporal.py:43-53) + optional max_pool1d(2) (temporal.py:137-139) in one
 * launch, 128 channels, tcgen05 TF32:  out = [pool]( relu_final( conv1x1(relu(conv_k3_dil(x) + bd)) + b1 + x ) ).
 * The intermediate activation stays in shared memory as the second GEMM's operand.  Wd_kco
 * [3][128][128], W1_kco [128][128] (Conv1d weights permuted to [k][Cout][Cin]).  tiles: device array
 * of {int64 row0; int64 row0_out; int32 t0; int32 T} (24 bytes): one entry per 128-row tile; row0_out
 * is the video's first row in `out` (the pooled resolution when pool != 0, else == row0). */
int mucon_wavenet_layer_tf32(const float* x, float* out, const float* Wd_kco, const float* bd,
                             const float* W1_kco, const float* b1, const void* tiles, int num_tiles,
                             int64_t rows, int dilation, int pool, int relu_final, void* stream);
/* The same layer with the weight traffic shared between the two CTAs of a thread-block cluster: each
 * CTA loads half of every weight k-block with a TMA multicast to both, so a tile reads 128 KB of
 * weights from L2 instead of 256 KB.  `tiles` must list the two tiles of a pair next to each other,
 * both from the same video: pad every video to an even number of tiles with {row0, row0_out,
 * t0 = 128 * tiles_of_video, T} entries (all rows of such a tile lie beyond T, nothing is stored);
 * num_tiles is then even. */
int mucon_wavenet_layer_tf32_pair(const float* x, float* out, const float* Wd_kco, const float* bd,
                                  const float* W1_kco, const float* b1, const void* tiles, int num_tiles,
                                  int64_t rows, int dilation, int pool, int relu_final, void* stream);
/* One whole WaveNet layer (temporal.py:43-53, + max_pool1d(2) :137-139, + the ReLU of :144) on tcgen05 kind::f16:
 * 16-bit activations ([rows, 128] row-major, 32-byte aligned) and weights (Wd_kco [3][128][128], W1_kco
 * [128][128], resident in shared memory for the whole launch), fp32 accumulate.  fp16 == 0: bfloat16; fp16 != 0:
 * IEEE half (11-bit mantissa: the rounding of the residual stream, which dominates the bf16 path's error, is 8x
 * smaller; values saturate at +-65504).  The two bias vectors are HOST arrays of 128 floats (they travel as a
 * kernel parameter).  `out` ([rows_out, 128]) has the activation type (pooled resolution when pool != 0) or, with
 * out_f32 != 0 (pool must be 0), is fp32.  tiles as for mucon_wavenet_layer_tf32. */
int mucon_wavenet_layer_bf16(const void* x, void* out, const void* Wd_kco, const float* bd_h, const void* W1_kco,
                             const float* b1_h, const void* tiles, int num_tiles, int64_t rows, int64_t rows_out,
                             int dilation, int pool, int relu_final, int out_f32, int fp16, void* stream);
/* The same launch with the skip connection optional: residual == 0 gives out = conv_1x1(relu(conv_k3(x) + bd)) + b1.
 * With an identity centre tap, bd = 0 and a dilation no video reaches (side taps then see only padding and are never
 * loaded) that is a plain 1x1 convolution of non-negative 16-bit rows on the resident-weight pipeline: how the fast
 * path runs `last_conv` (temporal.py:144-145, its input has just been through the ReLU of :144). */
int mucon_wavenet_layer_bf16_ex(const void* x, void* out, const void* Wd_kco, const float* bd_h, const void* W1_kco,
                                const float* b1_h, const void* tiles, int num_tiles, int64_t rows, int64_t rows_out,
                                int dilation, int pool, int relu_final, int out_f32, int fp16, int residual,
                                void* stream);
/* k = 1 or k = 3 dilated Conv1d with padding = dilation (temporal.py:21-31,48-52), fp32:
 *   out[t, co] = bias[co] + sum_tap sum_ci W_tco[tap][ci][co] * f(in[t + (tap - taps/2)*dilation, ci])
 * f = ReLU when relu_in; ReLU on the result when relu_out; `residual` ([rows, Cout] or NULL) is
 * added last (y += x, temporal.py:52).  W_tco is the Conv1d weight permuted to [tap][Cin][Cout]. */
int mucon_conv1d(const float* in, float* out, const float* W_tco, const float* bias, const float* residual,
                 const int64_t* row_off, int V, int max_T, int Cin, int Cout, int taps, int dilation,
                 int relu_in, int relu_out, void* stream);
/* max_pool1d(kernel_size=2) per video (temporal.py:139): off_out rows = floor(off_in rows / 2). */
int mucon_maxpool2(const float* in, float* out, const int64_t* off_in, const int64_t* off_out, int V,
                   int max_T_out, int C, void* stream);
/* GroupNorm(groups, C) over each video's (T x C/groups) elements + optional ReLU (models.py:759-764). */
int mucon_groupnorm_relu(const float* in, float* out, const float* gamma, const float* beta,
                         const int64_t* row_off, int V, int C, int groups, float eps, int relu, void* stream);
/* log_softmax over C classes of the pooled-resolution logits, expanded to every frame with the
 * nearest-neighbour index of F.interpolate: out[t,:] = lsm(logits[min(floor(t*(float)Tz/T), Tz-1), :])
 * (models.py:574-580 with the 1x1 classifier applied before the upsample, and :368). */
int mucon_logsoftmax_expand(const float* logits, const int64_t* off_z, const int64_t* off_t, int V, int max_T,
                            int C, float* out, void* stream);
/* log_softmax over the C classes of every row (same arithmetic as mucon_logsoftmax_expand, no expansion):
 * the pooled-resolution table mucon_viterbi_align_fused_pooled reads. */
int mucon_logsoftmax_rows(const float* logits, int64_t rows, int C, float* out, void* stream);
/* Fused tail at the pooled resolution, 128 hidden channels, up to 64 classes: GroupNorm(groups) over each video
 * (+ ReLU) (models.py:759-764), the 1x1 classifier (:276-278, :580; Wc_hc = weight permuted to [hidden][classes],
 * fp32 FFMA) and log_softmax (:368):  x [rows, 128] -> lsm_out [rows, classes] (and z_out [rows, 128] = the
 * normalised activations, or NULL).  tiles: the 16-byte {int64 row0; int32 t0; int32 T} records of
 * mucon_conv_gemm_tf32 at this resolution, tile_vid[num_tiles] the video of each tile, stats_ws: 2 * V * groups
 * floats of workspace.  Two launches (statistics, then everything else). */
int mucon_tail_logprobs(const float* x, const int64_t* row_off, const void* tiles, const int32_t* tile_vid,
                        int num_tiles, int V, int H, int groups, float eps, int relu, const float* gamma,
                        const float* beta, const float* Wc_hc, const float* bc, int num_classes, float* stats_ws,
                        float* z_out, float* lsm_out, void* stream);
/* out[t,:] = table[min(floor(t*(float)Tz/T), Tz-1), :] per video: the nearest-neighbour expansion of
 * F.interpolate (models.py:574-576) alone (the drop-in's materialising path). */
int mucon_expand_rows(const float* table, const int64_t* off_z, const int64_t* off_t, int V, int max_T, int C,
                      float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * vit_mof counters: nearest-neighbour resize of each video's predicted labels
 * to its ground-truth length (src/core/utils.py:34-47) + MoF counts (src/core/metrics/segmentation.py
 * :16-44, evaluators.py:225-243).  pred_off / gt_off: [V+1] offsets into pred / gt.  ignore_ids_h:
 * host array of up to 16 target ids to skip.  counts[2*v] = correct, counts[2*v+1] = total
 * (zeroed by the call); MoF = sum correct / sum total. */
int mucon_vit_mof(const int32_t* pred, const int64_t* pred_off, const int32_t* gt, const int64_t* gt_off,
                  int V, int max_T_gt, const int32_t* ignore_ids_h, int n_ignore,
                  unsigned long long* counts, void* stream);

/* Segment-level metrics of the Viterbi head, one CTA per video: the predicted labels are
 * resized to the ground truth's length (src/core/utils.py:34-47), both vectors are cut into maximal runs whose
 * label is not in ignore_ids, then IoD / IoU (src/core/metrics/isba_code.py:22-109), the normalised edit score and
 * the F1 matching counts at overlaps 0.1 / 0.25 / 0.5 (mstcn_code.py:27-81) as consumed at
 * src/mucon/evaluators.py:230-243.  out[v][12] = {iod, iou, edit, tp, fp, fn, tp, fp, fn, tp, fp, fn} (doubles;
 * iod / iou / edit are NaN where the reference divides by zero).  ws: 8-byte aligned workspace of
 * mucon_vit_segment_metrics_ws_words(sum of ground-truth frames, V) int32 words.  Only 96 bytes per video leave
 * the GPU. */
int mucon_vit_segment_metrics(const int32_t* pred, const int64_t* pred_off, const int32_t* gt, const int64_t* gt_off,
                              int V, const int32_t* ignore_ids_h, int n_ignore, int32_t* ws, double* out,
                              void* stream);
int64_t mucon_vit_segment_metrics_ws_words(int64_t total_gt_frames, int V);

#ifdef __cplusplus
}
#endif
#endif /* MUCON_B200_H_ */
