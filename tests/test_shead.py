"""The s-head at test time (reference src/mucon/models.py:585-745) against outputs of the UNMODIFIED reference model
(tests/golden/shead.npz, minted by tests/golden/make_golden_shead.py on a real MuCon built from the reference's default
configuration).  fp32 arithmetic in a different summation order than torch's LSTM / Linear kernels: log-probabilities
and length logits within 2e-4 absolute (measured ~1e-5); greedy tokens identical."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shead.npz"))
N_CASES = int(G["n"])


def _load(model):
    sd = {k[2:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("w.")}
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    return model


def test_state_dict_names_match_reference():
    from mucon_b200.shead import SHead
    m = SHead(num_classes=48)
    ref = {k[2:]: G[k].shape for k in G.files if k.startswith("w.")}
    own = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert own == {k: tuple(s) for k, s in ref.items()}


@pytest.mark.gpu
def test_shead_matches_reference(cuda_device):
    from mucon_b200.shead import SHead
    m = _load(SHead(num_classes=48)).to(cuda_device).eval()
    zs = [torch.from_numpy(G[f"c{i}_z"]) for i in range(N_CASES)]
    off_h = np.concatenate([[0], np.cumsum([z.shape[0] for z in zs])]).astype(np.int64)
    off = torch.from_numpy(off_h).to(cuda_device)
    z = torch.cat(zs).to(cuda_device)
    tf = [G[f"c{i}_tf_in"] for i in range(N_CASES)]
    out = m.forward_packed(z, off, off_h, tf, teacher_forcing=True)
    for i in range(N_CASES):
        n = tf[i].shape[0]
        assert int(out["n_steps"][i]) == n
        got_lp, got_len = out["logp"][i, :n].cpu().numpy(), out["lengths"][i, :n].cpu().numpy()
        assert np.abs(got_lp - G[f"c{i}_tf_logp"]).max() <= 2e-4, np.abs(got_lp - G[f"c{i}_tf_logp"]).max()
        assert np.abs(got_len - G[f"c{i}_tf_len"]).max() <= 2e-4, np.abs(got_len - G[f"c{i}_tf_len"]).max()
    out = m.forward_packed(z, off, off_h, None, teacher_forcing=False, max_steps=31)
    for i in range(N_CASES):
        want_lp = G[f"c{i}_greedy_logp"]
        n = want_lp.shape[0]
        assert int(out["n_steps"][i]) == n
        assert np.array_equal(out["tokens"][i, :n].cpu().numpy(), want_lp.argmax(1))
        assert np.abs(out["logp"][i, :n].cpu().numpy() - want_lp).max() <= 5e-4
        assert np.abs(out["lengths"][i, :n].cpu().numpy() - G[f"c{i}_greedy_len"]).max() <= 5e-4
    # the reference's one-video signature
    pt, pl = m.sequence_generation_forward(zs[1][None].to(cuda_device), tf[1].shape[0], torch.from_numpy(tf[1]))
    assert len(pt) == tf[1].shape[0] and pt[0].shape == (1, 49)
    assert np.abs(torch.cat(pt).cpu().numpy() - G["c1_tf_logp"]).max() <= 2e-4


@pytest.mark.gpu
def test_shead_greedy_stops_at_eos(cuda_device):
    """an s-head whose transcript head always prefers EOS stops after one step; n_steps and the padding say so"""
    from mucon_b200.shead import SHead
    torch.manual_seed(0)
    m = SHead(num_classes=48).to(cuda_device).eval()
    with torch.no_grad():
        m.fs_decoder_transcript[2].bias.zero_()
        m.fs_decoder_transcript[2].bias[48] = 50.0
    off_h = np.array([0, 20, 55], dtype=np.int64)
    z = torch.randn(55, 128, device=cuda_device).relu()
    out = m.forward_packed(z, torch.from_numpy(off_h).to(cuda_device), off_h, None, teacher_forcing=False)
    assert out["n_steps"].tolist() == [1, 1] and out["tokens"][:, 0].tolist() == [48, 48]
    assert (out["tokens"][:, 1:] == -1).all()


@pytest.mark.gpu
def test_full_inference_pipeline(cuda_device):
    """features -> backbone -> s-head (teacher-forced and greedy) -> predict -> class-mean lengths -> alignment from the
    pooled table: equal to running the evaluator glue on the host (class_mean_lengths + PoissonModel) over the expanded
    log-probabilities, video by video through the drop-in Viterbi class"""
    from mucon_b200 import PoissonModel, SingleTranscriptGrammar
    from mucon_b200.evaluate import class_mean_lengths
    from mucon_b200.inference import infer_and_align
    from mucon_b200.shead import SHead
    from mucon_b200.temporal import MuConBackbone
    from mucon_b200.viterbi import Viterbi, ViterbiEngine
    torch.manual_seed(2)
    rng = np.random.default_rng(2)
    net = MuConBackbone(input_feature_size=256).to(cuda_device).eval()
    sh = _load(SHead(num_classes=48)).to(cuda_device).eval()
    Ts = [900, 1700, 2400, 640]
    plan = net.plan(Ts, cuda_device)
    feats = torch.randn(int(sum(Ts)), 256, device=cuda_device).abs() * 0.5
    tf = [np.concatenate([[49], rng.integers(0, 48, n)]) for n in (4, 6, 7, 3)]
    eng = ViterbiEngine(cuda_device)
    for tfi in (tf, None):
        out = infer_and_align(net, sh, eng, feats, plan, 48, transcripts_tf_input=tfi)
        torch.cuda.synchronize()
        ap = out["plan"]
        res = eng.fetch(ap)
        lp = net.logprobs_packed(net.encode_packed(feats, plan), plan).cpu().numpy()
        dec = Viterbi(None, None, frame_sampling=30, device=cuda_device)
        off = np.concatenate([[0], np.cumsum(Ts)])
        for v in range(len(Ts)):
            tr = out["transcripts"][v]
            if tfi is not None:
                assert tr == [int(x) for x in tfi[v][1:]]
            if res["status"][v] != 0:
                continue   # K < N for a 30-token greedy transcript of random weights: -inf on both sides
            dec.grammar = SingleTranscriptGrammar(tr, 48)
            dec.length_model = PoissonModel(class_mean_lengths(tr, out["rel"][v].cpu().numpy(), Ts[v], 48))
            score, labels, segs = dec.decode(lp[off[v]:off[v + 1]])
            assert abs(score - res["score"][v]) <= 1e-9 * abs(score)
            assert labels == res["labels"][ap.vid_off[v]:ap.vid_off[v + 1]].tolist()


@pytest.mark.gpu
def test_shead_degenerate_lengths(cuda_device):
    """a video with no pooled rows (T < 16) next to normal ones, and an empty batch: finite outputs, no crash"""
    from mucon_b200.shead import SHead
    torch.manual_seed(1)
    m = SHead(num_classes=48).to(cuda_device).eval()
    off_h = np.array([0, 7, 7, 30], dtype=np.int64)          # the middle video has Tz = 0
    z = torch.randn(30, 128, device=cuda_device).relu()
    tf = [np.array([49, 3, 4]), np.array([49, 5]), np.array([49, 1, 2, 3])]
    out = m.forward_packed(z, torch.from_numpy(off_h).to(cuda_device), off_h, tf, teacher_forcing=True)
    assert out["n_steps"].tolist() == [3, 2, 4]
    for v, n in enumerate((3, 2, 4)):
        assert torch.isfinite(out["logp"][v, :n]).all() and torch.isfinite(out["lengths"][v, :n]).all()
    off0 = np.array([0], dtype=np.int64)
    out = m.forward_packed(torch.zeros(0, 128, device=cuda_device), torch.from_numpy(off0).to(cuda_device), off0, [],
                           teacher_forcing=True)
    assert out["logp"].shape[0] == 0
