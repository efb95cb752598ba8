"""Full test-time inference of the reference on the GPU for a packed batch of videos (BASELINE.json configs[1]):

    features -> backbone (temporal.MuConBackbone.infer_pooled_packed: projection, WaveNet layers, GroupNorm, classifier,
                log-softmax at the pooled resolution)                                 models.py:746-773, 567-582, 368
             -> s-head, greedy (shead.SHead.forward_packed)                          models.py:585-728
             -> MuCon.predict: transcript = argmax per step, relative lengths = softmax of the length logits of all
                steps but the last                                                    models.py:360-374
             -> evaluator glue: transcript without its last (EOS) entry, class-mean lengths, Poisson model
                (evaluate.class_mean_params_device)                                   evaluators.py:128-167
             -> transcript-constrained Viterbi alignment from the pooled table        evaluators.py:178, viterbi.py:49-158

One small device->host read in the middle (the predicted transcripts: the alignment plan's packing depends on them);
everything else stays on the device."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .evaluate import class_mean_params_device
from .viterbi import AlignPlan, default_seg0_f32


def predict_packed(backbone, shead, feats, plan, transcripts_tf_input=None, precision=None):
    """-> dict(table [sum Tz, C] log-probabilities, z_off, transcripts (list of int lists, EOS / last step dropped),
    rel (list of CUDA float32 tensors: relative lengths per transcript position), shead (raw s-head outputs))."""
    table, z_off, z = backbone.infer_pooled_packed(feats, plan, precision=precision, want_z=True)
    tf = transcripts_tf_input
    sh = shead.forward_packed(z, z_off, plan.off_host[-1], tf, teacher_forcing=tf is not None)
    n_steps = sh["n_steps"].cpu().numpy()
    tokens = sh["tokens"].cpu().numpy()
    transcripts, rel = [], []
    for v in range(plan.V):
        n = int(n_steps[v])
        if tf is not None:
            tr = [int(x) for x in np.asarray(tf[v])[1:n]]              # the teacher-forced target without EOS
        else:
            tr = [int(x) for x in tokens[v, :n - 1]]                    # models.py:364-366, evaluators.py:131
        transcripts.append(tr)
        rel.append(torch.softmax(sh["lengths"][v, :n - 1], dim=0))     # models.py:367 (lengths = all steps but the last)
    return dict(table=table, z_off=z_off, transcripts=transcripts, rel=rel, shead=sh)


def infer_and_align(backbone, shead, engine, feats, plan, n_classes, transcripts_tf_input=None, frame_sampling=30,
                    max_length=2000, precision=None):
    """predict_packed + the Viterbi block of the evaluator for the batch.  Videos whose predicted transcript is empty
    are aligned to the single label 0 (the reference would crash at one_hot, SURVEY V-edge).
    -> dict(plan, transcripts, rel, table, z_off); results are in plan.score / plan.labels / plan.seg_blocks."""
    p = predict_packed(backbone, shead, feats, plan, transcripts_tf_input, precision)
    trs = [tr if len(tr) else [0] for tr in p["transcripts"]]
    rel = [r if r.numel() else torch.ones(1, device=feats.device) for r in p["rel"]]
    T = plan.T[0]
    params = class_mean_params_device(rel, trs, T, feats.device)
    ap = AlignPlan(T, [[tr] for tr in trs], n_classes, fs=frame_sampling, max_len=max_length, len_params_dev=params,
                   device=feats.device, labels="best")
    try:
        engine.run(ap, p["table"], seg0_f32=default_seg0_f32(np.float32), write_bs=False, z_off=p["z_off"])
    except _lib.MuconError:
        # transcripts longer than the fused kernel's register budget (e.g. a 30-step greedy decode that never emits
        # EOS): expand the pooled table to every frame and take the two-kernel path (same results)
        lvl = len(plan.off) - 1
        full = torch.empty((plan.rows[0], p["table"].shape[1]), dtype=torch.float32, device=feats.device)
        _lib.check(_lib.lib().mucon_expand_rows(
            _lib.ptr(p["table"]), _lib.ptr(plan.off[lvl]), _lib.ptr(plan.off[0]), C.c_int(plan.V), C.c_int(plan.max_T[0]),
            C.c_int(full.shape[1]), _lib.ptr(full), C.c_void_p(torch.cuda.current_stream(feats.device).cuda_stream)),
            "mucon_expand_rows")
        engine.run(ap, full, seg0_f32=default_seg0_f32(np.float32))
    p["plan"] = ap
    p["transcripts"] = trs
    return p
