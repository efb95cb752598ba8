"""Mints tests/golden/evaluator_flow.npz by driving the UNMODIFIED reference evaluator
(/root/reference/src/mucon/evaluators.py: MuConEvaluator.on_start_eval / batch_eval_calculation /
on_finish_eval, with viterbi_mode(True)) over a few synthetic videos.  fandak / yacs / edit_distance are not
installed: they are stubbed with the minimum the call sites touch; np.float is aliased (removed from
NumPy).  The model is a stand-in whose predict() returns seeded y-head log-probabilities, s-head transcript and
relative lengths -- the evaluator only consumes MuConPredictOut.  What is frozen per video: the decode() inputs the
reference built (log-probs, transcript, class-mean lengths) and outputs (score, labels, segments), the resized
label vectors, and the evaluator's final vit_* metrics over all videos.  Re-run: python tests/golden/make_golden_evaluator_flow.py
"""
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
np.float = float  # removed alias used by core/metrics/isba_code.py:46 and mstcn_code.py:30
import scipy.signal  # noqa: E402
import scipy.signal.windows  # noqa: E402
scipy.signal.gaussian = scipy.signal.windows.gaussian  # moved in modern SciPy (mucon/masks.py:2)


class _Any(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (object,), {})


def install_stubs():
    for name in ("fandak", "fandak.utils", "fandak.utils.torch", "fandak.core", "fandak.core.datasets",
                 "fandak.core.evaluators", "fandak.core.trainers", "fandak.utils.misc", "yacs", "yacs.config",
                 "edit_distance"):
        sys.modules[name] = _Any(name)

    class Model(nn.Module):
        def __init__(self, cfg=None):
            super().__init__()
            self.cfg = cfg

    class Evaluator:
        def __init__(self, cfg, test_db, model, device):
            self.cfg, self.test_db, self.model, self.device = cfg, test_db, model, device

    sys.modules["fandak"].Model = Model
    sys.modules["fandak"].Evaluator = Evaluator
    sys.modules["fandak.utils.torch"].tensor_to_numpy = lambda t: t.detach().cpu().numpy()
    import difflib  # edit_distance.SequenceMatcher(a=, b=).ratio(): only feeds s_mat_score, not the hot path
    sys.modules["edit_distance"].SequenceMatcher = lambda a, b: difflib.SequenceMatcher(None, a, b)
    sys.path.insert(0, "/root/reference/src")


def main():
    install_stubs()
    from tests import synth
    from mucon.evaluators import MuConEvaluator
    from core.viterbi import viterbi as ref_viterbi

    NS = types.SimpleNamespace
    C = 48
    cfg = NS(evaluator=NS(viterbi=NS(multi_length=False)), system=NS(num_workers=0))
    db = NS(get_num_classes=lambda: C, background_class_ids=[0])

    class FakeModel:
        def set_teacher_forcing(self, mode):
            self.tf = mode

        def predict(self, batch, forward_out):
            return forward_out  # the test hands over a ready MuConPredictOut stand-in

    ev = MuConEvaluator(cfg, db, FakeModel(), "cpu")
    ev.viterbi_mode(True)
    ev.on_start_eval()

    # record what the reference hands to / gets from its own decoder (a wrapper around the unmodified method)
    rec = []
    orig_decode = ref_viterbi.Viterbi.decode

    def recording_decode(self, log_frame_probs):
        out = orig_decode(self, log_frame_probs)
        rec.append(dict(logp=np.array(log_frame_probs), transcript=list(self.grammar.transcript)
                        if hasattr(self.grammar, "transcript") else None,
                        means=np.array(self.length_model.mean_lengths), score=out[0], labels=np.array(out[1]),
                        segs=np.array([(s.label, s.length) for s in out[2]])))
        return out

    ref_viterbi.Viterbi.decode = recording_decode
    rng = np.random.default_rng(123)
    out = {}
    n_videos = 6
    for v in range(n_videos):
        N = int(rng.integers(2, 9))
        T = int(rng.integers(30 * N + 5, 2600))
        tr = [int(x) for x in rng.integers(0, C, N)]
        if v == 1:
            tr = [0, 5, 7, 5, 12, 0]
            N = 6
        logp, seg_len = synth.planted_logp(rng, T, C, tr, np.float32)
        planted = np.repeat(np.array(tr), seg_len)
        rel = torch.softmax(torch.from_numpy(rng.normal(size=N).astype(np.float32)), 0)
        # ground truth at another length (the evaluator resizes predictions to it): the planted segmentation, jittered
        Tg = int(T * rng.uniform(0.6, 1.5))
        gt = planted[np.minimum((np.arange(Tg) * T / Tg).astype(np.int64), T - 1)].copy()
        flip = rng.random(Tg) < 0.004
        gt[flip] = rng.integers(0, C, int(flip.sum()))
        batch = NS(feats=torch.zeros(1, T, 4), transcript=torch.tensor(tr), gt_label=torch.from_numpy(gt))
        pred = NS(transcript=tr + [C], lengths=rel, segmentation_logits=torch.from_numpy(logp))
        ev.batch_eval_calculation(batch, pred)
        r = rec[-1]
        assert r["transcript"] is None or r["transcript"] == tr
        out[f"v{v}_logp"], out[f"v{v}_tr"], out[f"v{v}_rel"] = logp, np.array(tr), rel.numpy()
        out[f"v{v}_gt"] = gt
        out[f"v{v}_means"], out[f"v{v}_score"] = r["means"], np.float64(r["score"])
        out[f"v{v}_labels"], out[f"v{v}_segs"] = r["labels"].astype(np.int32), r["segs"].astype(np.int64)
        out[f"v{v}_vit_resized"] = np.asarray(ev.vit_segs[-1]).astype(np.int32)
        out[f"v{v}_y_resized"] = np.asarray(ev.y_segs[-1]).astype(np.int32)
        # per-video values of the vit_* metric objects
        out[f"v{v}_iod"], out[f"v{v}_iou"] = np.float64(ev.vit_iod_metric.values[-1]), np.float64(ev.vit_iou_metric.values[-1])
        out[f"v{v}_iod_nbg"] = np.float64(ev.vit_iod_nbg_metric.values[-1])
        out[f"v{v}_iou_nbg"] = np.float64(ev.vit_iou_nbg_metric.values[-1])
        out[f"v{v}_edit"] = np.float64(ev.vit_edit_score_metric.values[-1])
    ref_viterbi.Viterbi.decode = orig_decode
    res = ev.on_finish_eval()
    for k in ("vit_mof", "vit_mof_nbg", "vit_iod", "vit_iou", "vit_iod_nbg", "vit_iou_nbg", "vit_edit_score", "y_mof"):
        out["final_" + k] = np.float64(getattr(res, k))
    out["final_vit_f1_score"] = np.array(res.vit_f1_score, dtype=np.float64)
    out["final_vit_f1_tp_fp_fn"] = np.array([ev.vit_f1_score_metric.tp, ev.vit_f1_score_metric.fp, ev.vit_f1_score_metric.fn])
    out["n"] = np.int64(n_videos)
    out["numpy_version"] = np.__version__
    np.savez_compressed(os.path.join(HERE, "evaluator_flow.npz"), **out)
    print("wrote evaluator_flow.npz:", {k: float(out[k]) for k in out if k.startswith("final_") and np.ndim(out[k]) == 0},
          out["final_vit_f1_score"])


if __name__ == "__main__":
    main()
