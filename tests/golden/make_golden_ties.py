"""Mints tests/golden/ties.npz: candidate sets whose best final hypotheses TIE exactly (duplicated, integer-valued
class columns), decoded by the UNMODIFIED reference (/root/reference/src/core/viterbi) with ModifiedPathGrammar.
Which candidate the reference returns then depends only on the order of its insertion-ordered hypothesis dict
(viterbi.py:26-28,93-138) -- the rule mucon_viterbi_select_ranked / grammar.tie_ranks restate.
Re-run with:  python tests/golden/make_golden_ties.py"""
import os
import random
import sys

import numpy as np

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from core.viterbi.grammar import ModifiedPathGrammar, SingleTranscriptGrammar  # noqa: E402
from core.viterbi.length_model import PoissonModel  # noqa: E402
from core.viterbi.viterbi import Viterbi  # noqa: E402

N_CASES = 160


def main():
    rng = random.Random(7)
    nrng = np.random.default_rng(7)
    out = {}
    kept = naive_wrong = 0
    while kept < N_CASES:
        C = rng.choice([4, 8, 12, 20])
        fs = rng.choice([1, 2, 3])
        T = rng.randrange(6, 30) * fs + rng.randrange(0, fs)
        trs = set()
        ntr = rng.randrange(2, 7)
        while len(trs) < ntr:
            trs.add(tuple(rng.randrange(C) for _ in range(rng.randrange(1, 5))))
        trs = sorted(trs)
        rng.shuffle(trs)
        base = nrng.integers(-3, 0, size=(T, 2)).astype(np.float64)
        logp = np.ascontiguousarray(base[:, nrng.integers(0, 2, size=C)])
        if rng.random() < 0.5:
            means = np.full(C, float(rng.choice([2, 3, 5])) * fs)
        else:
            means = nrng.integers(1, 4, size=C).astype(np.float64) * fs
        max_len = 40 * fs
        lm = PoissonModel(means, max_length=max_len)
        score, labels, segs = Viterbi(ModifiedPathGrammar([list(t) for t in trs], C), lm, frame_sampling=fs).decode(logp)
        if not np.isfinite(score):
            continue
        # how many candidates reach the winning score on their own (>= 2: a real tie), and would "lowest index" differ?
        singles = [Viterbi(SingleTranscriptGrammar(list(t), C), lm, frame_sampling=fs).decode(logp) for t in trs]
        tied = [i for i, s in enumerate(singles) if s[0] == score]
        if len(tied) < 2:
            continue
        naive_wrong += list(singles[tied[0]][1]) != list(labels)
        k = f"t{kept}_"
        out[k + "logp"] = logp
        out[k + "means"] = means
        out[k + "fs_maxlen_C"] = np.array([fs, max_len, C], dtype=np.int64)
        out[k + "tr_len"] = np.array([len(t) for t in trs], dtype=np.int64)
        out[k + "tr"] = np.concatenate([np.asarray(t, dtype=np.int32) for t in trs])
        out[k + "score"] = np.float64(score)
        out[k + "labels"] = np.asarray(labels, dtype=np.int32)
        out[k + "seg_label"] = np.asarray([s.label for s in segs], dtype=np.int32)
        out[k + "seg_length"] = np.asarray([s.length for s in segs], dtype=np.int64)
        kept += 1
    out["n_cases"] = np.int64(kept)
    np.savez_compressed(os.path.join(HERE, "ties.npz"), **out)
    print(f"{kept} tied cases, 'lowest index wins' would return different labels in {naive_wrong}")


if __name__ == "__main__":
    main()
