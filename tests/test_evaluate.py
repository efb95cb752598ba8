"""The evaluator's Viterbi block (reference src/mucon/evaluators.py:147-180) for a batch."""
import os

import numpy as np
import pytest

from tests import synth
from tests.util import load_golden, same_score

HERE = os.path.dirname(os.path.abspath(__file__))


def test_class_mean_lengths_match_reference_block():
    """tests/golden/eval_lengths.npz holds outputs of the reference's own statements."""
    from mucon_b200.evaluate import class_mean_lengths
    g = np.load(os.path.join(HERE, "golden", "eval_lengths.npz"))
    for i in range(int(g["n"])):
        got = class_mean_lengths(g[f"tr{i}"], g[f"rel{i}"], int(g[f"T{i}"]), int(g[f"C{i}"]))
        assert got.dtype == np.float64
        assert np.array_equal(got, g[f"lengths{i}"]), i


@pytest.mark.gpu
def test_align_videos_equals_per_video_drop_in(cuda_device):
    import torch
    from mucon_b200 import PoissonModel, SingleTranscriptGrammar
    from mucon_b200.evaluate import align_videos, class_mean_lengths
    from mucon_b200.viterbi import Viterbi, ViterbiEngine
    from oracle import metrics as ometrics
    rng = np.random.default_rng(11)
    C = 48
    logps, trs, rels, gts = [], [], [], []
    for i in range(12):
        N = int(rng.integers(1, 10))
        T = int(rng.integers(max(60, 30 * N), 4000))
        tr = list(map(int, rng.integers(0, C, N)))
        lp, _ = synth.planted_logp(rng, T, C, tr, np.float32)
        logps.append(lp)
        trs.append(tr)
        rels.append(rng.dirichlet(3 * np.ones(N)).astype(np.float32))
        gts.append(rng.integers(0, C, int(rng.integers(T // 2, 2 * T))).astype(np.int32))
    eng = ViterbiEngine(cuda_device)
    res = align_videos(eng, logps, trs, rels, C, targets=gts, ignore_ids=(0,))
    dec = Viterbi(None, None, frame_sampling=30, device=cuda_device)
    correct = total = 0
    for v in range(len(logps)):
        dec.grammar = SingleTranscriptGrammar(trs[v], C)
        dec.length_model = PoissonModel(class_mean_lengths(trs[v], rels[v], logps[v].shape[0], C))
        score, labels, segs = dec.decode(logps[v])
        assert same_score(res["score"][v], score)
        assert res["labels"][v].tolist() == labels
        assert res["segments"][v] == [(s.label, s.length) for s in segs]
        c, t = ometrics.mof_counts(gts[v], ometrics.same_size_interpolate(labels, len(gts[v])), ignore_ids=(0,))
        assert res["mof_counts"][v].tolist() == [c, t]
        correct += c
        total += t
    assert res["mof"] == pytest.approx(correct / max(total, 1))
    # the golden c1 video through the batched entry point
    g = load_golden("c1_f32")
    rel = np.full(len(g["transcripts"][0]), 1.0 / len(g["transcripts"][0]), dtype=np.float32)
    r1 = align_videos(eng, (torch.from_numpy(g["logp"]).to(cuda_device), [g["logp"].shape[0]]),
                      [g["transcripts"][0]], [rel], 48)
    assert len(r1["labels"][0]) == 2000


@pytest.mark.gpu
def test_align_videos_errors(cuda_device):
    from mucon_b200.evaluate import align_videos
    from mucon_b200.viterbi import ViterbiEngine
    eng = ViterbiEngine(cuda_device)
    with pytest.raises(AttributeError):  # K > N*J
        align_videos(eng, [np.zeros((3990, 8), np.float32)], [[0, 1]], [np.array([0.5, 0.5], np.float32)], 8)
    with pytest.raises(IndexError):      # T < frame_sampling
        align_videos(eng, [np.zeros((20, 8), np.float32)], [[0]], [np.array([1.0], np.float32)], 8)
