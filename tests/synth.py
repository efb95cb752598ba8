"""Seeded synthetic Breakfast-shaped inputs shared by tests, smoke() and bench.py
(SURVEY.md section 8d).  Pure NumPy; no reference, no oracle, no CUDA."""
import numpy as np


def planted_logp(rng, T, C, transcript, dtype=np.float32, boost=3.0):
    """log_softmax(N(0,1) + boost * onehot(planted segmentation)) -- [T, C]."""
    N = len(transcript)
    cuts = np.sort(rng.choice(np.arange(1, T), size=N - 1, replace=False)) if N > 1 else np.array([], dtype=int)
    bounds = np.concatenate([[0], cuts, [T]])
    x = rng.standard_normal((T, C))
    for n in range(N):
        x[bounds[n]:bounds[n + 1], transcript[n]] += boost
    x = x - x.max(axis=1, keepdims=True)
    x = x - np.log(np.exp(x).sum(axis=1, keepdims=True))
    return x.astype(dtype), np.diff(bounds)


def class_means(rel, transcript, C, T, floor=1.0):
    """Per-class mean absolute lengths the evaluator feeds to PoissonModel
    (reference src/mucon/evaluators.py:155-165), zeros -> 1.  Means are floored at `floor`
    because the reference's PoissonModel turns any mean < 0.5 into NaN scores (SURVEY.md V-edge)."""
    rel = np.asarray(rel, dtype=np.float32)
    tr = np.asarray(transcript)
    tot = np.zeros(C, dtype=np.float64)
    cnt = np.zeros(C, dtype=np.float64)
    np.add.at(tot, tr, rel.astype(np.float64))
    np.add.at(cnt, tr, 1.0)
    tot *= T
    cnt[cnt == 0] = 1
    tot /= cnt
    tot[tot == 0] = 1
    return np.maximum(tot, floor)


def breakfast_split(seed=0, V=1712, C=48, fs=30, J=66, t_lo=300, t_hi=10000, n_hi=12):
    """c2: V videos, T ~ clip(lognormal(ln 1800, 0.7)), N ~ U{max(2, ceil(K/J)) .. min(n_hi, K)}."""
    rng = np.random.default_rng(seed)
    T = np.clip(np.round(rng.lognormal(np.log(1800.0), 0.7, V)), t_lo, t_hi).astype(np.int64)
    K = T // fs
    lo = np.maximum(2, -(-K // J))
    hi = np.minimum(n_hi, K)
    N = rng.integers(lo, hi + 1)
    transcripts = [rng.integers(0, C, int(n)).astype(np.int32) for n in N]
    return T, transcripts


def random_edits(rng, tr, C, n_cands, n_lo, n_hi):
    """c3: candidate 0 is the transcript itself, the rest are random insert/delete/substitute edits."""
    out = [list(map(int, tr))]
    while len(out) < n_cands:
        c = list(out[0])
        for _ in range(int(rng.integers(1, 4))):
            op = int(rng.integers(0, 3))
            if op == 0 and len(c) < n_hi:
                c.insert(int(rng.integers(0, len(c) + 1)), int(rng.integers(0, C)))
            elif op == 1 and len(c) > max(1, n_lo):
                c.pop(int(rng.integers(0, len(c))))
            else:
                c[int(rng.integers(0, len(c)))] = int(rng.integers(0, C))
        if n_lo <= len(c) <= n_hi:
            out.append(c)
    return out
