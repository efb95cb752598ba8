"""Mints tests/golden/poisson_f32.npz: the UNMODIFIED reference PoissonModel (/root/reference/src/core/viterbi/
length_model.py:42-83) built from FLOAT32 mean lengths -- its norms and l * log(m) - m then run in float32 and are promoted
by NumPy's rules -- and one reference Viterbi.decode with that model."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/src")
from core.viterbi.grammar import SingleTranscriptGrammar  # noqa: E402
from core.viterbi.length_model import PoissonModel  # noqa: E402
from core.viterbi.viterbi import Viterbi  # noqa: E402
from tests import synth  # noqa: E402


def main():
    rng = np.random.default_rng(21)
    C, T = 48, 1500
    tr = [3, 17, 3, 40, 0]
    means = synth.class_means(rng.dirichlet(5 * np.ones(5)).astype(np.float32), tr, C, T).astype(np.float32)
    lm = PoissonModel(means)
    logp, _ = synth.planted_logp(rng, T, C, tr, np.float32)
    dec = Viterbi(SingleTranscriptGrammar(tr, C), lm, frame_sampling=30)
    score, labels, segs = dec.decode(logp)
    out = dict(means=means, table_rows=lm.poisson[np.arange(1, 67) * 30], table_head=lm.poisson[:40], transcript=np.array(tr),
               logp=logp, score=np.float64(score), labels=np.array(labels, dtype=np.int32),
               segs=np.array([(s.label, s.length) for s in segs]), numpy_version=np.__version__)
    np.savez_compressed(os.path.join(HERE, "poisson_f32.npz"), **out)
    print("wrote poisson_f32.npz", score)


if __name__ == "__main__":
    main()
