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


@pytest.mark.gpu
def test_device_class_mean_params_match_host(cuda_device):
    """mucon_class_mean_params (class means + Poisson parameters on the device) against the host statements pinned by
    eval_lengths.npz: m exactly, ln m and the norms to 2 ulp (ln() is CUDA's); and the batched alignment built on them
    gives the host path's labels and segments, scores to 1e-12 relative."""
    import torch
    from mucon_b200.evaluate import align_videos, class_mean_lengths, class_mean_params_device
    from mucon_b200.length_model import poisson_params
    from mucon_b200.viterbi import ViterbiEngine
    g = np.load(os.path.join(HERE, "golden", "eval_lengths.npz"))
    n = int(g["n"])
    trs = [g[f"tr{i}"].astype(np.int64).tolist() for i in range(n)]
    rels = [g[f"rel{i}"].astype(np.float32) for i in range(n)]
    Ts = [int(g[f"T{i}"]) for i in range(n)]
    got = class_mean_params_device(rels, trs, Ts, cuda_device).cpu().numpy()
    pos = 0
    for i in range(n):
        want = poisson_params(g[f"lengths{i}"])[np.asarray(trs[i])]
        blk = got[pos:pos + len(trs[i])]
        pos += len(trs[i])
        assert np.allclose(blk[:, 1], want[:, 1], rtol=4e-16, atol=0), i          # the means (float64 sums)
        assert np.allclose(blk[:, 0], want[:, 0], rtol=5e-16, atol=1e-15), i
        ok = np.isfinite(want[:, 2])
        assert np.allclose(blk[ok, 2], want[ok, 2], rtol=1e-13, atol=1e-10), i
    rng = np.random.default_rng(21)
    C = 48
    logps, trs, rels = [], [], []
    for i in range(40):
        N = int(rng.integers(1, 10))
        T = int(rng.integers(max(60, 30 * N), min(4000, 30 * 66 * N)))
        tr = list(map(int, rng.integers(0, C, N)))
        lp, _ = synth.planted_logp(rng, T, C, tr, np.float32)
        logps.append(lp)
        trs.append(tr)
        rels.append(rng.dirichlet(3 * np.ones(N)).astype(np.float32))
    eng = ViterbiEngine(cuda_device)
    a = align_videos(eng, logps, trs, rels, C)
    b = align_videos(eng, logps, trs, rels, C, device_lengths=True)
    assert np.allclose(a["score"], b["score"], rtol=1e-12, atol=0)
    assert a["segments"] == b["segments"]
    assert all(np.array_equal(x, y) for x, y in zip(a["labels"], b["labels"]))
