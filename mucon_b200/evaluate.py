"""The evaluator's Viterbi block for a whole batch of videos (reference
src/mucon/evaluators.py:147-180, one video at a time there): class-mean lengths from the s-head's
relative lengths, a Poisson length model per video, one alignment launch for the batch, and --
optionally -- the vit_mof counters (evaluators.py:225-243) without the labels leaving the GPU.

    lengths = class_mean_lengths(transcript, relative_lengths, n_frames, n_classes)   # :155-165
    out = align_videos(engine, log_probs, transcripts, relative_lengths, n_classes)   # :148-180
"""
import numpy as np
import torch

from . import _lib
from .length_model import poisson_params
from .viterbi import AlignPlan, ViterbiEngine, default_seg0_f32


def class_mean_lengths(transcript, relative_lengths, n_frames, n_classes):
    """evaluators.py:155-165: absolute mean length of every class from the lengths the s head
    predicted for the actions of the transcript; classes that do not occur get 1.

    transcript [N] ints, relative_lengths [N] (float32 softmax output in the reference) -> float64 [C].
    Same operations in the same order and dtypes: float32 rel-lengths . float64 one-hot -> float64,
    times n_frames, divided by the occurrence count, exact zeros replaced by 1."""
    tr = np.asarray(transcript).reshape(-1)
    actions = np.eye(int(n_classes))[tr]                      # one_hot, evaluators.py:71-72
    lengths = np.dot(np.asarray(relative_lengths), actions)
    lengths *= n_frames
    k = actions.sum(0)
    k[k == 0] = 1
    lengths /= k
    lengths[lengths == 0] = 1
    return lengths


def class_mean_params_device(relative_lengths, transcripts, T, device):
    """class_mean_lengths + PoissonModel parameters for a batch ON THE DEVICE (mucon_class_mean_params): returns the
    [sum N, 3] float64 CUDA tensor AlignPlan(len_params_dev=...) takes.  relative_lengths: per video [N_v] (CUDA /
    CPU tensors or arrays; float32 like the s-head's softmax output) or one concatenated CUDA tensor."""
    import ctypes as C
    from .length_model import _log_tail
    T = np.asarray(T, dtype=np.int64)
    V = int(T.shape[0])
    nlen = np.array([len(tr) for tr in transcripts], dtype=np.int64)
    tr_off = np.concatenate([[0], np.cumsum(nlen)]).astype(np.int32)
    tr = np.concatenate([np.asarray(t, dtype=np.int32) for t in transcripts]) if V else np.zeros(0, np.int32)
    if torch.is_tensor(relative_lengths):
        rel = relative_lengths.to(device=device, dtype=torch.float32).contiguous()
    else:
        rel = torch.cat([r.detach().to(device).float() if torch.is_tensor(r) else torch.from_numpy(np.asarray(r, np.float32)).to(device)
                         for r in relative_lengths]).contiguous() if V else torch.zeros(0, device=device)
    n_tail = int(T.max(initial=1)) + 2
    tail = torch.from_numpy(np.ascontiguousarray(_log_tail(n_tail)[:n_tail + 1])).to(device)
    meta = torch.from_numpy(np.concatenate([np.concatenate([[0], np.cumsum(T)]).astype(np.int64).view(np.int32),
                                            tr_off, tr])).to(device)
    vid_off = meta[:2 * (V + 1)].view(torch.int64)
    tr_off_d, tr_d = meta[2 * (V + 1):2 * (V + 1) + V + 1], meta[2 * (V + 1) + V + 1:]
    out = torch.empty((int(tr_off[-1]), 3), dtype=torch.float64, device=device)
    _lib.check(_lib.lib().mucon_class_mean_params(
        _lib.ptr(rel), _lib.ptr(tr_d), _lib.ptr(tr_off_d), _lib.ptr(vid_off), C.c_int(V), _lib.ptr(tail),
        C.c_int(int(tail.numel())), _lib.ptr(out), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
        "mucon_class_mean_params")
    return out


def align_videos(engine, log_probs, transcripts, relative_lengths, n_classes, frame_sampling=30,
                 max_length=2000, np_mode=None, targets=None, ignore_ids=(), device_lengths=False):
    """Batched form of the evaluator's decode step.

    log_probs: list of [T_v, C] arrays/tensors (float32/float64), or a tuple (packed, T) of an
    already concatenated CUDA tensor [sum T, C] and the list of lengths.
    transcripts: per video, the predicted transcript (list of ints, no EOS).
    relative_lengths: per video, [N_v] relative lengths (they sum to 1).
    targets: optional per-video int target labels (any length; resized like make_same_size_interpolate)
    -> dict(score [V] float64, labels list of int32 arrays, segments list of (label, length) lists,
            plan, and with targets: mof_counts int64 [V, 2], mof float).
    """
    if engine is None:
        engine = ViterbiEngine()
    dev = engine.device
    if isinstance(log_probs, tuple):
        packed, T = log_probs
        T = np.asarray(T, dtype=np.int64)
        if not packed.is_cuda:
            packed = packed.to(dev)
    else:
        T = np.asarray([int(x.shape[0]) for x in log_probs], dtype=np.int64)
        arrs = [x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x)) for x in log_probs]
        packed = torch.cat([a.to(dev) for a in arrs]) if arrs else torch.zeros((0, n_classes), device=dev)
    packed = packed.contiguous()
    V = int(T.shape[0])
    if len(transcripts) != V or len(relative_lengths) != V:
        raise ValueError("one transcript and one relative-length vector per video")
    fs = int(frame_sampling)
    if V and int(T.min()) < fs:
        raise IndexError(f"a sequence is shorter than frame_sampling={fs}")  # viterbi.py:87
    if device_lengths:
        # class means and Poisson parameters computed on the device (no per-video host loop; ln() is CUDA's, so path
        # scores can differ from the default host path in the last bits -- labels only on exact near-ties)
        means = None
        plan = AlignPlan(T, [[list(map(int, tr))] for tr in transcripts], n_classes, fs=fs, max_len=int(max_length),
                         len_params_dev=class_mean_params_device(relative_lengths, transcripts, T, dev), device=dev,
                         labels="best")
    else:
        means = np.stack([class_mean_lengths(tr, _np(rl), int(t), n_classes)
                          for tr, rl, t in zip(transcripts, relative_lengths, T)]) if V else np.zeros((0, n_classes))
        plan = AlignPlan(T, [[list(map(int, tr))] for tr in transcripts], n_classes, fs=fs, max_len=int(max_length),
                         len_params=poisson_params(means), device=dev, labels="best")
    is32 = packed.dtype == torch.float32
    if np_mode is None:
        seg0 = default_seg0_f32(np.float32 if is32 else np.float64)
    else:
        seg0 = is32 and np_mode == "numpy2"
    engine.run(plan, packed, seg0_f32=seg0, write_bs=False)
    res = {"plan": plan, "means": means}
    if targets is not None:
        from .metrics import mof, mof_counts
        gt = torch.cat([torch.as_tensor(np.asarray(g), dtype=torch.int32) for g in targets]).to(dev)
        gt_off = np.concatenate([[0], np.cumsum([len(g) for g in targets])])
        counts = mof_counts(plan.labels, plan.vid_off, gt, gt_off, ignore_ids)
        res["mof_counts"] = counts.cpu()
        res["mof"] = mof(res["mof_counts"])
    out = engine.fetch(plan)
    bad = np.nonzero(out["status"] == _lib.UNIT_INFEASIBLE)[0]
    if bad.size:
        # the reference dies in traceback when every hypothesis has been dropped (K > N*J)
        raise AttributeError(f"no hypothesis survives for video(s) {bad.tolist()}: too long for the transcript")
    res["score"] = out["score"]
    res["labels"] = [out["labels"][plan.vid_off[v]:plan.vid_off[v + 1]] for v in range(V)]
    segs = []
    for v in range(V):
        sb = out["seg_blocks"][plan.tr_off[v]:plan.tr_off[v + 1]]
        tr = transcripts[v]
        s = [[int(tr[n]), int(fs * sb[n])] for n in range(len(tr)) if sb[n] > 0]
        s[-1][1] += int(T[v]) - fs * int(T[v] // fs)
        segs.append([tuple(x) for x in s])
    res["segments"] = segs
    return res


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
