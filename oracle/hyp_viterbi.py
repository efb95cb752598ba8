"""Hypothesis-table Viterbi decoder -- ORACLE / TEST INFRASTRUCTURE ONLY.

A from-scratch CPU restatement of the reference decoder's algorithm
(reference: src/core/viterbi/viterbi.py:49-158, grammar semantics of
src/core/viterbi/grammar.py:143-217, length-model semantics of
src/core/viterbi/length_model.py:76-83).  It keeps the reference's data flow --
an insertion-ordered table of hypotheses keyed by (label history, current
label, current length) with linked traceback records -- so that its cost
profile on the host is the one the reference has; bench.py times *this* as the
"reference CPU path" on the GPU box, where /root/reference does not exist.

Arithmetic follows the reference expression by expression so that NumPy's
scalar promotion (NEP 50 in NumPy >= 2) produces the same float32/float64 mix:
  start   : 0.0 + F[fs-1, l]                              (viterbi.py:85-88)
  stay    : s + bs(t, l)                                  (viterbi.py:99)
  advance : s + bs(t, l_old) + len(length, l_old) + 0.0   (viterbi.py:111-116)
  final   : s + len(length, l) + g_end                    (viterbi.py:130-134)
with bs(t, l) = F[t, l] - F[t-fs, l] (F[t, l] if t < fs)  (viterbi.py:68-72).

Inputs are lowered: a grammar is a list of candidate transcripts (a single
transcript == SingleTranscriptGrammar, several == (Modified)PathGrammar) and
the length model is a table ``len_table[length, label]`` plus ``max_len``
(score is -inf for length >= max_len, length_model.py:76-80).
"""
import numpy as np

START, END = -1, -2  # grammar.py:19-24


def build_successors(candidates):
    """Prefix tree: history tuple -> set of admissible next labels (grammar.py:201-207).  Built with the reference's
    own statement ({x}.union(old)): the ITERATION order of these sets fixes the order of the hypothesis table and
    with it which of several equally-scored final hypotheses wins (tests/golden/ties.npz)."""
    succ = {}
    for tr in candidates:
        path = [int(x) for x in tr] + [END]
        for i, nxt in enumerate(path):
            hist = (START,) + tuple(path[:i])
            succ[hist] = {nxt}.union(succ.get(hist, set()))
    return succ


def _put(table, key, score, rec):
    # last writer wins on ties: replace iff old <= new (viterbi.py:26-28)
    old = table.get(key)
    if old is None or old[0] <= score:
        table[key] = (score, rec)


def decode(logp, candidates, len_table, max_len=2000, fs=30):
    """Returns (score, labels list[int] of len T, segments list[(label, length)]).

    Raises the same way the reference does on infeasible inputs (AttributeError /
    IndexError family) -- callers in tests only check that *an* exception is raised.
    """
    logp = np.asarray(logp)
    T = logp.shape[0]
    succ = build_successors(candidates)
    F = np.cumsum(logp, axis=0)  # viterbi.py:51 -- sequential, dtype of logp

    def bs(t, l):
        return F[t, l] - F[t - fs, l] if t >= fs else F[t, l]

    def lscore(length, l):
        return -np.inf if length >= max_len else len_table[length, l]

    def gscore(hist, l):
        return 0.0 if l in succ.get(hist, ()) else -np.inf

    # records are (label, predecessor record, boundary flag)  (viterbi.py:13-17)
    table = {}
    root = (START,)
    for l in succ.get(root, ()):
        _put(table, root + (l, fs), gscore(root, l) + bs(fs - 1, l), (l, None, True))

    for t in range(2 * fs - 1, T, fs):  # viterbi.py:57-59
        nxt = {}
        for key, (s, rec) in table.items():
            hist, l, length = key[:-2], key[-2], key[-1]
            if length + fs <= max_len:
                _put(nxt, hist + (l, length + fs), s + bs(t, l), (l, rec, False))
            hist2 = hist + (l,)
            for l2 in succ.get(hist2, ()):
                if l2 == END:
                    continue
                sc = s + bs(t, l) + lscore(length, l) + gscore(hist2, l2)
                _put(nxt, hist2 + (l2, fs), sc, (l2, rec, True))
        table = nxt

    best_s, best_rec = -np.inf, None
    for key, (s, rec) in table.items():
        hist, l, length = key[:-2], key[-2], key[-1]
        sc = s + lscore(length, l) + gscore(hist + (l,), END)
        if sc >= best_s:
            best_s, best_rec = sc, rec

    # traceback (viterbi.py:140-158): fs frames per record; the T - fs*K leftover
    # frames take the final record's label and end up FIRST after the reversal.
    rec = best_rec
    labels = []
    segs = [[rec[0], 0]]  # AttributeError-equivalent (TypeError) if best_rec is None
    while rec is not None:
        segs[-1][1] += fs
        labels += [rec[0]] * fs
        if rec[2] and rec[1] is not None:
            segs.append([rec[1][0], 0])
        rec = rec[1]
    segs[0][1] += T - len(labels)
    labels += [best_rec[0]] * (T - len(labels))
    labels.reverse()
    segs.reverse()
    return best_s, labels, [(int(a), int(b)) for a, b in segs]
