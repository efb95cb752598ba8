"""Dense (K x N x J) restatement of the transcript-constrained Viterbi -- ORACLE ONLY.

Equivalent to the reference decoder (src/core/viterbi/viterbi.py:49-158) for a
single transcript (SingleTranscriptGrammar, grammar.py:196-217); candidate sets
(PathGrammar, grammar.py:143-191) decode as the arg-max over per-candidate runs
because prefix-tree keys never merge.

State S[n][j], j = 1..J (J = max_len // fs): best score of "in segment n for j
blocks".  Block k covers frames [fs*k, fs*(k+1)).  Per step k >= 1:
    a        = S[n][j] + bs_k[tr_n]                    (stay, viterbi.py:97-104)
    S'[n][j+1] = a                       if j < J
    cand_j   = (a + rows[n][j]) + 0.0                  (advance, viterbi.py:106-121;
                                                        block k is scored with the OLD label)
    S'[n+1][1] = fold_j cand_j  with "replace iff old <= new", j ascending
                                                       (viterbi.py:26-28)  -> bp[k][n+1] = j
Final: fold_j (S[N-1][j] + rows[N-1][j]) + 0.0, same fold            (viterbi.py:125-138)
Traceback: T - fs*K leftover frames go FIRST with the LAST label     (viterbi.py:154-158)

Dtypes: F and bs carry the dtype of logp; with float32
log-probs under NumPy >= 2 segment 0 accumulates in float32 (``seg0_f32``),
everything else is float64.
"""
import numpy as np


class Infeasible(Exception):
    """K > N*J (all hypotheses die) or T < fs -- the reference raises here too."""


def numpy_seg0_f32(logp_dtype):
    """The float mix the installed NumPy gives the reference code."""
    return np.dtype(logp_dtype) == np.float32 and int(np.__version__.split(".")[0]) >= 2


def block_scores(logp, fs):
    """bs[k, c] in logp's dtype, from the sequential cumulative sum (viterbi.py:51,68-72)."""
    T = logp.shape[0]
    K = T // fs
    F = np.cumsum(logp, axis=0)
    ends = F[fs - 1 : fs * K : fs]  # rows fs*(k+1)-1
    bs = ends.copy()
    bs[1:] = ends[1:] - ends[:-1]
    return bs


def length_rows(len_table, transcript, fs, max_len):
    """rows[n, j-1] = length score of j blocks in label tr_n; -inf when j*fs >= max_len."""
    J = max_len // fs
    rows = np.full((len(transcript), J), -np.inf, dtype=np.float64)
    for j in range(1, J + 1):
        if j * fs < max_len:
            rows[:, j - 1] = len_table[j * fs, list(transcript)]
    return rows


def _fold(cands, js):
    """Sequential 'replace iff old <= new' over candidates in ascending j."""
    best, bj = cands[0], js[0]
    for c, j in zip(cands[1:], js[1:]):
        if best <= c:
            best, bj = c, j
    return best, bj


def decode(logp, transcript, rows, fs=30, seg0_f32=None):
    """Returns dict(score, labels[T] int32, seg_blocks[N] int32, bp[K,N] uint16, jf)."""
    logp = np.asarray(logp)
    if seg0_f32 is None:
        seg0_f32 = numpy_seg0_f32(logp.dtype)
    T = logp.shape[0]
    N, J = rows.shape
    K = T // fs
    if K < 1 or N < 1 or K > N * J:
        raise Infeasible(f"T={T} fs={fs} N={N} J={J}")
    bs = block_scores(logp, fs)
    tr = list(transcript)

    S = np.zeros((N, J + 1), dtype=np.float64)  # index j = 1..J
    live = np.zeros((N, J + 1), dtype=bool)
    bp = np.zeros((K, N), dtype=np.uint16)
    first = np.float32(0.0 + bs[0, tr[0]]) if seg0_f32 else np.float64(0.0 + bs[0, tr[0]])
    S[0, 1], live[0, 1] = first, True

    for k in range(1, K):
        S2 = np.zeros_like(S)
        live2 = np.zeros_like(live)
        for n in range(N):
            b = bs[k, tr[n]]
            js = np.nonzero(live[n])[0]
            if js.size == 0:
                continue
            if n == 0 and seg0_f32:
                a = (S[n, js].astype(np.float32) + np.float32(b)).astype(np.float64)
            else:
                a = S[n, js] + np.float64(b)
            keep = js < J
            S2[n, js[keep] + 1] = a[keep]
            live2[n, js[keep] + 1] = True
            if n + 1 < N:
                cand = (a + rows[n, js - 1]) + 0.0
                best, bj = _fold(list(cand), list(js))
                S2[n + 1, 1], live2[n + 1, 1] = best, True
                bp[k, n + 1] = bj
        S, live = S2, live2

    seg_blocks = np.zeros(N, dtype=np.int32)
    if K < N:
        # nothing reaches the last segment: every final score is -inf and the fold
        # keeps the last-inserted hypothesis, which is (segment K-1, 1 block);
        # its history is one block per segment.
        score, jf = -np.inf, 1
        seg_blocks[:K] = 1
        last = K - 1
    else:
        js = np.nonzero(live[N - 1])[0]
        cand = (S[N - 1, js] + rows[N - 1, js - 1]) + 0.0
        score, jf = _fold(list(cand), list(js))
        n, k0 = N - 1, K - int(jf)
        seg_blocks[n] = jf
        while n > 0:
            ln = int(bp[k0, n])
            seg_blocks[n - 1] = ln
            k0 -= ln
            n -= 1
        assert k0 == 0
        last = N - 1
    rem = T - fs * K
    labels = np.empty(T, dtype=np.int32)
    labels[:rem] = tr[last]
    pos = rem
    for n in range(last + 1):
        labels[pos : pos + fs * seg_blocks[n]] = tr[n]
        pos += fs * seg_blocks[n]
    assert pos == T
    return dict(score=np.float64(score), labels=labels, seg_blocks=seg_blocks, bp=bp, jf=int(jf))


def segments_from_blocks(seg_blocks, transcript, fs, T):
    """[(label, length_in_frames)] the way the reference reports them (viterbi.py:141-158)."""
    K = T // fs
    segs = [(int(transcript[n]), int(fs * b)) for n, b in enumerate(seg_blocks) if b > 0]
    lab, ln = segs[-1]
    segs[-1] = (lab, ln + T - fs * K)
    return segs
