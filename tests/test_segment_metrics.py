"""IoD / IoU / Edit / F1 of the Viterbi head (reference src/core/metrics/isba_code.py:22-109, mstcn_code.py:27-81,
consumed at src/mucon/evaluators.py:230-243): the NumPy restatement against values the unmodified reference evaluator
produced (tests/golden/evaluator_flow.npz), and the device kernel against both."""
import os

import numpy as np
import pytest

from oracle import metrics as om

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "evaluator_flow.npz"))
OVERLAPS = (0.1, 0.25, 0.5)


def test_oracle_segment_metrics_match_reference_evaluator():
    tp, fp, fn = np.zeros(3), np.zeros(3), np.zeros(3)
    for v in range(int(G["n"])):
        P, Y = G[f"v{v}_vit_resized"], G[f"v{v}_gt"]
        assert om.iod(P, Y) == float(G[f"v{v}_iod"])
        assert om.iou(P, Y) == float(G[f"v{v}_iou"])
        assert om.iod(P, Y, (0,)) == float(G[f"v{v}_iod_nbg"])
        assert om.iou(P, Y, (0,)) == float(G[f"v{v}_iou_nbg"])
        assert om.edit_score(P, Y) == float(G[f"v{v}_edit"])
        for k, o in enumerate(OVERLAPS):
            a, b, c = om.f_counts(P, Y, o)
            tp[k] += a; fp[k] += b; fn[k] += c
    assert np.array_equal(np.array([tp, fp, fn]), G["final_vit_f1_tp_fp_fn"])
    assert [om.f1(*x) for x in zip(tp, fp, fn)] == G["final_vit_f1_score"].tolist()


def test_oracle_resize_and_mof_match_reference_evaluator():
    c = t = cn = tn = 0
    for v in range(int(G["n"])):
        r = om.same_size_interpolate(G[f"v{v}_labels"], len(G[f"v{v}_gt"]))
        assert np.array_equal(r, G[f"v{v}_vit_resized"])
        a, b = om.mof_counts(G[f"v{v}_gt"], r)
        c += a; t += b
        a, b = om.mof_counts(G[f"v{v}_gt"], r, (0,))
        cn += a; tn += b
    assert c / t == float(G["final_vit_mof"]) and cn / tn == float(G["final_vit_mof_nbg"])


@pytest.mark.gpu
@pytest.mark.parametrize("ignore", [(), (0,), (0, 3)])
def test_device_segment_metrics(cuda_device, ignore):
    """mucon_vit_segment_metrics on the golden videos plus random label vectors (few and many segments, vectors that
    are all background, single-frame videos): integer counters exact, ratios to 1e-12."""
    import torch
    from mucon_b200 import metrics as mm
    rng = np.random.default_rng(5)
    preds = [G[f"v{v}_labels"] for v in range(int(G["n"]))]
    gts = [G[f"v{v}_gt"] for v in range(int(G["n"]))]
    for i in range(20):
        Tp, Tg = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        nseg = int(rng.integers(1, 40))
        def seq(T):
            cuts = np.sort(rng.integers(0, T, nseg))
            return rng.integers(0, 6, nseg + 1)[np.searchsorted(cuts, np.arange(T), side="right")].astype(np.int32)
        preds.append(seq(Tp)); gts.append(seq(Tg))
    preds.append(np.zeros(50, np.int32)); gts.append(np.zeros(70, np.int32))          # all background
    preds.append(rng.integers(0, 48, 400).astype(np.int32)); gts.append(rng.integers(0, 48, 333).astype(np.int32))
    po = np.concatenate([[0], np.cumsum([len(p) for p in preds])])
    go = np.concatenate([[0], np.cumsum([len(g) for g in gts])])
    out = mm.segment_metrics(torch.from_numpy(np.concatenate(preds).astype(np.int32)).to(cuda_device), po,
                             torch.from_numpy(np.concatenate(gts).astype(np.int32)).to(cuda_device), go,
                             ignore_ids=ignore).cpu().numpy()
    for v, (p, g) in enumerate(zip(preds, gts)):
        P = om.same_size_interpolate(p, len(g))
        want = [om.iod(P, g, ignore), om.iou(P, g, ignore), om.edit_score(P, g, ignore)]
        for o in OVERLAPS:
            want += list(om.f_counts(P, g, o, ignore))
        got = out[v]
        for k in range(3):
            assert (np.isnan(want[k]) and np.isnan(got[k])) or abs(got[k] - want[k]) <= 1e-12 * max(1.0, abs(want[k])), (v, k, got[k], want[k])
        assert got[3:].tolist() == want[3:], (v, got[3:], want[3:])


@pytest.mark.gpu
def test_evaluator_flow_with_gpu_pieces(cuda_device):
    """The reference evaluator's Viterbi block (evaluators.py:147-180, 225-243) replayed with the GPU pieces swapped
    in -- drop-in Viterbi / SingleTranscriptGrammar / PoissonModel classes fed exactly what the reference built (its
    recorded decode() inputs), then the on-device resize + MoF + IoD / IoU / Edit / F1 -- against what the unmodified
    reference evaluator returned for the same videos."""
    import torch
    from mucon_b200 import PoissonModel, SingleTranscriptGrammar
    from mucon_b200 import metrics as mm
    from mucon_b200.evaluate import class_mean_lengths
    from mucon_b200.viterbi import Viterbi
    dec = Viterbi(None, None, frame_sampling=30, device=cuda_device)
    n = int(G["n"])
    labels = []
    for v in range(n):
        tr = G[f"v{v}_tr"].tolist()
        lengths = class_mean_lengths(tr, G[f"v{v}_rel"], G[f"v{v}_logp"].shape[0], 48)
        assert np.array_equal(lengths, G[f"v{v}_means"])
        dec.grammar = SingleTranscriptGrammar(tr, 48)
        dec.length_model = PoissonModel(lengths)
        score, lab, segs = dec.decode(G[f"v{v}_logp"])
        assert float(score) == float(G[f"v{v}_score"])
        assert np.array_equal(np.asarray(lab, dtype=np.int32), G[f"v{v}_labels"])
        assert [(s.label, s.length) for s in segs] == [tuple(x) for x in G[f"v{v}_segs"].tolist()]
        labels.append(np.asarray(lab, dtype=np.int32))
    gts = [G[f"v{v}_gt"].astype(np.int32) for v in range(n)]
    po = np.concatenate([[0], np.cumsum([len(p) for p in labels])])
    go = np.concatenate([[0], np.cumsum([len(g) for g in gts])])
    pd = torch.from_numpy(np.concatenate(labels)).to(cuda_device)
    gd = torch.from_numpy(np.concatenate(gts)).to(cuda_device)
    assert mm.mof(mm.mof_counts(pd, po, gd, go)) == float(G["final_vit_mof"])
    assert mm.mof(mm.mof_counts(pd, po, gd, go, ignore_ids=(0,))) == float(G["final_vit_mof_nbg"])
    s = mm.summarize(mm.segment_metrics(pd, po, gd, go))
    sn = mm.summarize(mm.segment_metrics(pd, po, gd, go, ignore_ids=(0,)))
    assert s["iod"] == pytest.approx(float(G["final_vit_iod"]), rel=1e-12)
    assert s["iou"] == pytest.approx(float(G["final_vit_iou"]), rel=1e-12)
    assert sn["iod"] == pytest.approx(float(G["final_vit_iod_nbg"]), rel=1e-12)
    assert sn["iou"] == pytest.approx(float(G["final_vit_iou_nbg"]), rel=1e-12)
    assert s["edit"] == pytest.approx(float(G["final_vit_edit_score"]), rel=1e-12)
    assert s["f1"] == pytest.approx(G["final_vit_f1_score"].tolist(), rel=1e-12)
