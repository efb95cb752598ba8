"""Mints tests/golden/viterbi_*.npz by running the UNMODIFIED reference decoder
(/root/reference/src/core/viterbi) in the build container.  The reference has no tests or golden
vectors for this path, so these files are the pin: inputs + the reference's
score / labels / segments / back-pointers.  Re-run with:  python tests/golden/make_golden.py

Back-pointers are read off the reference's own traceback records: for every hypothesis that has
just entered a segment, the number of records back to (and including) the previous boundary is
the predecessor segment's length in blocks.
"""
import os
import sys

import numpy as np

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from core.viterbi.grammar import ModifiedPathGrammar, SingleTranscriptGrammar  # noqa: E402
from core.viterbi.length_model import PoissonModel  # noqa: E402
from core.viterbi.viterbi import Viterbi  # noqa: E402

from tests import synth  # noqa: E402


def run_reference(logp, transcripts, means, fs, max_len):
    C = logp.shape[1]
    single = len(transcripts) == 1
    gr = SingleTranscriptGrammar(list(map(int, transcripts[0])), C) if single else \
        ModifiedPathGrammar([list(map(int, t)) for t in transcripts], C)
    dec = Viterbi(gr, PoissonModel(np.asarray(means, dtype=np.float64), max_length=max_len), frame_sampling=fs)
    K = logp.shape[0] // fs
    N = len(transcripts[0])
    bp = np.zeros((K, N), dtype=np.uint16)
    step = [0]
    orig = dec.decode_frame

    def spy(t, old, fsc):
        new = orig(t, old, fsc)
        step[0] += 1
        if single:
            for key, hyp in new.items():
                if key[-1] == fs:  # just entered segment n
                    n = len(key) - 3
                    node, cnt = hyp.traceback.predecessor, 0
                    while node is not None:
                        cnt += 1
                        if node.boundary:
                            break
                        node = node.predecessor
                    bp[step[0], n] = cnt
        return new

    dec.decode_frame = spy
    score, labels, segs = dec.decode(logp)
    return dict(score=np.float64(score), labels=np.asarray(labels, dtype=np.int32),
                seg_label=np.asarray([s.label for s in segs], dtype=np.int32),
                seg_length=np.asarray([s.length for s in segs], dtype=np.int64), bp=bp)


def pack_transcripts(trs):
    off = np.concatenate([[0], np.cumsum([len(t) for t in trs])]).astype(np.int32)
    return np.concatenate([np.asarray(t, dtype=np.int32) for t in trs]), off


def case(name, logp, transcripts, means, fs=30, max_len=2000):
    out = run_reference(logp, transcripts, means, fs, max_len)
    flat, off = pack_transcripts(transcripts)
    np.savez_compressed(os.path.join(HERE, f"viterbi_{name}.npz"), logp=logp, tr=flat, tr_off=off,
                        means=np.asarray(means, dtype=np.float64), fs=fs, max_len=max_len,
                        numpy_version=np.__version__, **out)
    print(f"{name:28s} T={logp.shape[0]:5d} C={logp.shape[1]:3d} cands={len(transcripts):2d} "
          f"score={out['score']:.6f} segs={list(zip(out['seg_label'], out['seg_length']))[:4]}")


def main():
    rng = np.random.default_rng(20260101)
    # 1. remainder quirk: T=100, tr=[1,2,3] -> 10 leftover frames FIRST, labelled 3
    logp = np.log(rng.dirichlet(np.ones(5), 100)).astype(np.float32)
    case("remainder_quirk", logp, [[1, 2, 3]], np.full(5, 33.0))
    # 2. all ties: constant log-probs, equal means -> largest predecessor length wins (V5)
    case("all_ties", np.full((600, 4), -1.25, dtype=np.float32), [[0, 1, 2, 3]], np.full(4, 150.0))
    case("all_ties_f64", np.full((600, 4), -1.25, dtype=np.float64), [[0, 1, 2, 3]], np.full(4, 150.0))
    # 3. integer-valued log-probs (exact arithmetic, many ties)
    li = -rng.integers(0, 3, (900, 6)).astype(np.float32)
    case("integer_valued", li, [[2, 0, 5, 1]], np.array([200.0, 250.0, 180.0, 1.0, 1.0, 270.0]))
    # 4. c1: Breakfast-shaped single video with repeated labels
    tr = [0, 5, 7, 5, 12, 0]
    lp, seglen = synth.planted_logp(rng, 2000, 48, tr, np.float32)
    rel = rng.dirichlet(5 * np.ones(6)).astype(np.float32)
    means = synth.class_means(rel, tr, 48, 2000)
    case("c1_f32", lp, [tr], means)
    case("c1_f64", lp.astype(np.float64), [tr], means)
    # 5. K == J*N exactly: T=3960, N=2 -> both segments 1980 frames
    lp = np.log(rng.dirichlet(np.ones(4), 3960)).astype(np.float32)
    case("k_eq_jn", lp, [[1, 2]], np.array([1.0, 1900.0, 2100.0, 1.0]))
    # 6. K < N: nothing reaches the last segment, score -inf, partial path (V7)
    lp = np.log(rng.dirichlet(np.ones(6), 100)).astype(np.float32)
    case("k_lt_n", lp, [[0, 1, 2, 3, 4, 5]], np.full(6, 16.0))
    # 7. other sampling rates / length caps (max_len % fs == 0 makes the top length score -inf)
    lp = np.log(rng.dirichlet(np.ones(7), 333)).astype(np.float32)
    case("fs7_len91", lp, [[3, 1, 4, 1, 5]], rng.uniform(20, 120, 7), fs=7, max_len=91)
    lp = np.log(rng.dirichlet(np.ones(5), 57)).astype(np.float64)
    case("fs1_len20", lp, [[0, 2, 4, 1]], rng.uniform(5, 25, 5), fs=1, max_len=20)
    # 8. candidate set (ModifiedPathGrammar): result == best single-transcript decode (V10)
    base = [4, 9, 2, 7]
    lp, _ = synth.planted_logp(rng, 1500, 12, base, np.float32)
    cands = synth.random_edits(np.random.default_rng(5), base, 12, 6, 2, 8)
    case("path_grammar_6", lp, cands, synth.class_means(rng.dirichlet(np.ones(4)).astype(np.float32), base, 12, 1500))
    # 9. a long video: T=10000, N=12
    tr = list(map(int, rng.integers(0, 16, 12)))
    lp, _ = synth.planted_logp(rng, 10000, 16, tr, np.float32)
    case("long_T10000_N12", lp, [tr], synth.class_means(rng.dirichlet(np.ones(12)).astype(np.float32), tr, 16, 10000))
    # 10. shapes beyond the register-resident kernels (generic kernel): J = 200, J = 300 (uint16
    #     back-pointers; frame_sampling = 1 is the reference class's default), N = 70
    tr = [3, 0, 2, 5, 1]
    lp, _ = synth.planted_logp(rng, 700, 6, tr, np.float32)
    case("generic_fs2_J200", lp, [tr], rng.uniform(60, 220, 6), fs=2, max_len=400)
    tr = [1, 3, 0, 2]
    lp, _ = synth.planted_logp(rng, 500, 4, tr, np.float64)
    case("generic_fs1_J300", lp, [tr], rng.uniform(60, 200, 4), fs=1, max_len=300)
    tr = list(map(int, rng.integers(0, 20, 70)))
    lp, _ = synth.planted_logp(rng, 4230, 20, tr, np.float32)
    case("generic_N70", lp, [tr], rng.uniform(30, 120, 20))


if __name__ == "__main__":
    main()
