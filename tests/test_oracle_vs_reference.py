"""CPU, build container only: the oracle restatements against the LIVE reference implementation
imported unmodified from /root/reference/src (skipped where that tree does not exist)."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        from core.viterbi.grammar import ModifiedPathGrammar, SingleTranscriptGrammar
        from core.viterbi.length_model import PoissonModel
        from core.viterbi.viterbi import Viterbi
        yield dict(Viterbi=Viterbi, Single=SingleTranscriptGrammar, Path=ModifiedPathGrammar, Poisson=PoissonModel)
    finally:
        sys.path.remove(REF)


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    for trial in range(n):
        C = int(rng.integers(3, 14))
        fs = int(rng.choice([30, 30, 7, 1, 13]))
        max_len = int(rng.choice([2000, 200, 91, 60])) if fs > 1 else int(rng.choice([20, 35]))
        J = max_len // fs
        N = int(rng.integers(1, 6))
        K = int(rng.integers(max(1, N - 1), N * J + 1))
        T = K * fs + int(rng.integers(0, fs))
        dt = [np.float32, np.float64][trial % 2]
        mode = trial % 4
        if mode == 0:
            logp = rng.standard_normal((T, C))
        elif mode == 1:
            logp = np.full((T, C), -1.5)
        elif mode == 2:
            logp = -rng.integers(0, 4, (T, C)).astype(float)
        else:
            logp = np.log(rng.dirichlet(np.ones(C), T))
        tr = [int(x) for x in rng.integers(0, C, N)]
        means = rng.uniform(0.6, T, C) if mode != 1 else np.full(C, T / N)
        yield logp.astype(dt), tr, means, fs, max_len


def test_poisson_table_bit_exact(ref):
    from oracle import poisson
    rng = np.random.default_rng(0)
    for max_len in (2000, 91, 20):
        means = np.concatenate([rng.uniform(0.5, 6000, 20), [1.0, 0.7, 2.5, 3.5, 1999.5]])
        pm = ref["Poisson"](means, max_length=max_len)
        assert np.array_equal(poisson.poisson_table(means, max_len), pm.poisson)
        assert np.array_equal(poisson.poisson_params(means)[:, 2], pm.norms)


def test_product_poisson_model_bit_exact(ref):
    from mucon_b200.length_model import PoissonModel
    rng = np.random.default_rng(1)
    means = rng.uniform(0.5, 9000, 48)
    a, b = PoissonModel(means), ref["Poisson"](means)
    assert np.array_equal(a.poisson, b.poisson) and np.array_equal(a.norms, b.norms)
    for l, c in [(30, 0), (1980, 5), (1999, 47), (2000, 3), (660, 11)]:
        assert a.score(l, c) == b.score(l, c)


def test_decoders_bit_exact_on_random_inputs(ref):
    from oracle import coracle, dense_viterbi, hyp_viterbi, poisson
    n_ok = 0
    for logp, tr, means, fs, max_len in _cases(60, 7):
        C = logp.shape[1]
        dec = ref["Viterbi"](ref["Single"](tr, C), ref["Poisson"](means, max_length=max_len), frame_sampling=fs)
        try:
            s, labels, segs = dec.decode(logp)
        except Exception:
            continue
        n_ok += 1
        tab = poisson.poisson_table(means, max_len)
        s2, l2, g2 = hyp_viterbi.decode(logp, [tr], tab, max_len, fs)
        assert (s == s2) and labels == l2 and [(x.label, x.length) for x in segs] == g2
        rows = dense_viterbi.length_rows(tab, tr, fs, max_len)
        d = dense_viterbi.decode(logp, tr, rows, fs)
        assert d["score"] == s and d["labels"].tolist() == labels
        c = coracle.decode_video(logp, tr, rows, fs, dense_viterbi.numpy_seg0_f32(logp.dtype))
        assert c["score"] == s and c["labels"].tolist() == labels
    assert n_ok >= 40


def test_candidate_set_equals_best_single(ref):
    from oracle import dense_viterbi, hyp_viterbi, poisson
    from tests import synth
    rng = np.random.default_rng(11)
    base = [3, 8, 1, 6, 2]
    logp, _ = synth.planted_logp(rng, 900, 10, base, np.float32)
    cands = synth.random_edits(rng, base, 10, 5, 2, 8)
    means = rng.uniform(50, 400, 10)
    dec = ref["Viterbi"](ref["Path"](cands, 10), ref["Poisson"](means), frame_sampling=30)
    s, labels, _ = dec.decode(logp)
    tab = poisson.poisson_table(means)
    s2, l2, _ = hyp_viterbi.decode(logp, cands, tab)
    assert s == s2 and labels == l2
    singles = [dense_viterbi.decode(logp, tr, dense_viterbi.length_rows(tab, tr, 30, 2000)) for tr in cands]
    best = max(singles, key=lambda d: d["score"])
    assert best["score"] == s and best["labels"].tolist() == labels
