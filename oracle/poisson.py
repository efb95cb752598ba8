"""Poisson length model -- ORACLE / TEST INFRASTRUCTURE ONLY.

Restates reference src/core/viterbi/length_model.py:42-83 (PoissonModel) and the
evaluator's class-mean glue src/mucon/evaluators.py:155-165.

  norms_c  = r*ln(r) - r - sum_{k=2..int(m_c)} ln k,  r = round-half-even(m_c)   (:54-63)
  tab[l,c] = ((l*ln(m_c) - m_c) - lf_l) - norms_c,    lf_l = sum_{i=1..l} ln i  (:65-71)
  tab[0,:] = -inf; score(l,c) = -inf for l >= max_len                           (:66, :76-80)

All sums are sequential float64 (np.cumsum is sequential), matching the
reference's Python accumulation loops bit for bit (checked in
tests/test_oracle_vs_reference.py).
"""
import numpy as np


def log_factorial_prefix(n):
    """lf[i] = sum_{k=1..i} ln k, sequential float64; lf[0] = 0."""
    out = np.zeros(n + 1, dtype=np.float64)
    if n >= 1:
        out[1:] = np.cumsum(np.log(np.arange(1, n + 1, dtype=np.float64)))
    return out


def poisson_params(means):
    """Per-class (ln m, m, norms) float64 triples, shape [C, 3]."""
    m = np.asarray(means, dtype=np.float64)
    r = np.round(m)
    with np.errstate(divide="ignore", invalid="ignore"):
        norms = r * np.log(r) - r
        top = int(np.max(m.astype(np.int64), initial=1))
        # sum_{k=2..i} ln k, sequential from k=2 (the reference starts its loop at 2)
        tail = np.zeros(max(top, 1) + 1, dtype=np.float64)
        if top >= 2:
            tail[2:] = np.cumsum(np.log(np.arange(2, top + 1, dtype=np.float64)))
        norms = norms - tail[np.maximum(m.astype(np.int64), 0)]
        logm = np.log(m)
    return np.stack([logm, m, norms], axis=1)


def poisson_table(means, max_len=2000):
    """Full [max_len, C] float64 table, row 0 = -inf (length_model.py:50,66-71)."""
    p = poisson_params(means)
    lf = log_factorial_prefix(max_len - 1)
    L = np.arange(max_len, dtype=np.float64)[:, None]
    with np.errstate(invalid="ignore"):
        tab = ((L * p[None, :, 0] - p[None, :, 1]) - lf[:, None]) - p[None, :, 2]
    tab[0, :] = -np.inf
    return tab


def class_mean_lengths(rel_lengths, transcript, n_classes, T):
    """evaluators.py:155-165: per-class mean absolute length, zeros -> 1.

    rel_lengths float32 [N] (softmax output), transcript list[int] len N.
    """
    actions = np.eye(n_classes)[np.asarray(transcript).reshape(-1)]  # one_hot, evaluators.py:71-72
    lengths = np.dot(np.asarray(rel_lengths), actions)
    lengths *= T
    k = actions.sum(0)
    k[k == 0] = 1
    lengths /= k
    lengths[lengths == 0] = 1
    return lengths


def poisson_table_loop(means, max_len=2000):
    """The table built the way the reference builds it -- a Python loop over all lengths
    (length_model.py:54-71).  Same values as poisson_table; used where the CPU baseline has to
    carry the reference's cost profile (bench.py)."""
    m = np.asarray(means, dtype=np.float64)
    tab = np.zeros((max_len, m.shape[0]))
    norms = np.round(m) * np.log(np.round(m)) - np.round(m)
    for c in range(len(m)):
        acc = 0
        for k in range(2, int(m[c]) + 1):
            acc += np.log(k)
        norms[c] = norms[c] - acc
    tab[0, :] = -np.inf
    acc = 0
    for l in range(1, max_len):
        acc += np.log(l)
        tab[l, :] = l * np.log(m) - m - acc - norms
    return tab
