"""The mutual-consistency loss around the GPU mask kernel.

Mirrors reference src/mucon/models.py:414-450 (`MuCon.mucon_loss`) and :452-525
(`calculate_mucon_loss_using_masks`, "flint" and "arithmetic" variants): softmax-projected lengths ->
`create_masks` (CUDA, autograd-aware, scales its argument in place like the reference) -> masked
class evidence -> NLL.  Everything after the masks is the same handful of torch ops the reference
uses, so gradients w.r.t. both the lengths and the frame logits flow as they do there.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .masks import TEMPLATES, create_masks, project_lengths_softmax

FLINT_CHUNK = 512  # kFlintChunk in csrc/masks.cu


def _flint_meta(Ms, Ts, device):
    """Offset tables for the fused evidence kernels (one small H2D copy)."""
    Ms, Ts = np.asarray(Ms, dtype=np.int64), np.asarray(Ts, dtype=np.int64)
    n_off = np.concatenate([[0], np.cumsum(Ms)]).astype(np.int32)
    seg_off = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
    row_vid = np.repeat(np.arange(Ms.shape[0], dtype=np.int32), Ms)
    nch = (Ts + FLINT_CHUNK - 1) // FLINT_CHUNK
    cv = np.repeat(np.arange(Ms.shape[0], dtype=np.int32), nch)
    first = np.concatenate([[0], np.cumsum(nch)])[:-1]
    ct0 = ((np.arange(int(nch.sum())) - np.repeat(first, nch)) * FLINT_CHUNK).astype(np.int32)
    chunks = np.stack([cv, ct0], 1).astype(np.int32).reshape(-1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return dict(n_off=dev(n_off), T=dev(Ts.astype(np.int32)), seg_off=dev(seg_off[:-1].copy()), row_vid=dev(row_vid),
                chunks=dev(chunks) if chunks.size else torch.zeros(2, dtype=torch.int32, device=device),
                n_chunks=int(nch.sum()), V=int(Ms.shape[0]), n_rows=int(n_off[-1]), max_rows=int(Ms.max(initial=0)),
                total_T=int(seg_off[-1]))


class _EvidenceFn(torch.autograd.Function):
    """E[r, c] = sum_t mask_r[t] * seg[t, c]  (mucon_flint_fwd / mucon_flint_bwd)."""

    @staticmethod
    def forward(ctx, L, seg, meta, overlap, tid, align):
        Lc = L.detach().float().contiguous()
        sg = seg.detach().float().contiguous()
        Cn = sg.shape[1]
        E = torch.empty((meta["n_rows"], Cn), dtype=torch.float32, device=sg.device)
        st = C.c_void_p(torch.cuda.current_stream(sg.device).cuda_stream)
        lib = _lib.lib()
        key = ("fwd_ws", Cn)
        if key not in meta:   # partial sums + row counters; zeroed once, the kernel leaves the counters at zero
            meta[key] = torch.zeros(int(lib.mucon_flint_fwd_ws_words(C.c_int(meta["n_rows"]), C.c_int(Cn))) + 4,
                                    dtype=torch.float32, device=sg.device)
        _lib.check(lib.mucon_flint_fwd_ws(
            _lib.ptr(Lc), _lib.ptr(meta["n_off"]), _lib.ptr(meta["T"]), _lib.ptr(meta["seg_off"]), _lib.ptr(meta["row_vid"]),
            C.c_int(meta["V"]), C.c_int(meta["n_rows"]), C.c_int(Cn), C.c_float(overlap), C.c_int(tid), C.c_int(align),
            _lib.ptr(sg), _lib.ptr(meta[key]), _lib.ptr(E), st), "mucon_flint_fwd_ws")
        ctx.save_for_backward(Lc, sg)
        ctx.meta, ctx.args = meta, (overlap, tid, align)
        return E

    @staticmethod
    def backward(ctx, gE):
        Lc, sg = ctx.saved_tensors
        meta = ctx.meta
        overlap, tid, align = ctx.args
        gE = gE.contiguous().float()
        Cn = sg.shape[1]
        gseg = torch.empty_like(sg) if ctx.needs_input_grad[1] else None
        ws = torch.empty(2 * meta["n_rows"], dtype=torch.float32, device=sg.device)
        gL = torch.empty(meta["n_rows"], dtype=torch.float32, device=sg.device)
        st = C.c_void_p(torch.cuda.current_stream(sg.device).cuda_stream)
        _lib.check(_lib.lib().mucon_flint_bwd(
            _lib.ptr(Lc), _lib.ptr(meta["n_off"]), _lib.ptr(meta["T"]), _lib.ptr(meta["seg_off"]), _lib.ptr(meta["row_vid"]),
            C.c_int(meta["V"]), C.c_int(meta["n_rows"]), C.c_int(meta["max_rows"]), C.c_int(Cn), C.c_float(overlap),
            C.c_int(tid), C.c_int(align), _lib.ptr(sg), _lib.ptr(gE), _lib.ptr(meta["chunks"]), C.c_int(meta["n_chunks"]),
            _lib.ptr(gseg), _lib.ptr(ws), _lib.ptr(gL), st), "mucon_flint_bwd")
        return gL, gseg, None, None, None, None


def flint_evidence(L, seg, Ms, Ts, overlap=0.0, template="box", align_corners=None, meta=None):
    """Masked class evidence of a batch without materialising the masks: L concatenated absolute
    lengths [sum Ms] (unscaled), seg packed frame logits [sum Ts, C] -> E [sum Ms, C].
    Differentiable w.r.t. both.  Needs C % 4 == 0, C <= 128, at most 64 segments per video."""
    if template not in TEMPLATES:
        raise NameError(f"Invalid template name ({template})")
    if not (L.is_cuda and seg.is_cuda):
        raise _lib.MuconError("flint_evidence needs CUDA tensors (there is no CPU fallback)")
    meta = meta if meta is not None else _flint_meta(Ms, Ts, seg.device)
    if seg.shape[0] != meta["total_T"] or L.shape[0] != meta["n_rows"]:
        raise ValueError("L / seg do not match Ms / Ts")
    return _EvidenceFn.apply(L, seg, meta, float(overlap), TEMPLATES[template], int(bool(align_corners)))


def flint_fusable(segmentation, n_segments):
    return (segmentation.is_cuda and segmentation.dtype == torch.float32 and segmentation.shape[1] % 4 == 0
            and segmentation.shape[1] <= 128 and n_segments <= 64)


def loss_from_masks(absolute_lengths, masks, segmentation, target_transcript, mucon_type="flint", class_weight=None):
    """models.py:452-525.  masks [N, T], segmentation [T, C] logits, target_transcript [N] long."""
    if mucon_type == "flint":
        # p_i = log_softmax((mask_i . seg) / L_i)  (models.py:459-468), all segments at once
        evidence = (masks @ segmentation) / absolute_lengths[:, None]
        return F.nll_loss(F.log_softmax(evidence, dim=1), target_transcript, weight=class_weight, reduction="mean")
    if mucon_type == "arithmetic":
        # sum_i sum_t CE(seg_t, tr_i) * mask_i[t] / T  (models.py:489-523)
        logp = F.log_softmax(segmentation, dim=1)
        ce = -logp[:, target_transcript]  # [T, N]
        if class_weight is not None:
            ce = ce * class_weight[target_transcript][None, :]
        return (ce * masks.t()).sum() / segmentation.size(0)
    raise Exception(f"Invalid mucon type ({mucon_type})")


def mucon_loss(lengths, segmentation, target_transcript, template="box", overlap=0.0, mucon_type="flint",
               class_weight=None, align_corners=None, fused=None):
    """models.py:414-450 given the s-head length logits [N], the frame logits [T, C] and the target
    transcript [N].  fused: None = use the fused evidence kernel when it applies (flint, float32
    CUDA logits, C % 4 == 0), False = materialise the masks like the reference."""
    T = segmentation.shape[0]
    absolute_lengths = project_lengths_softmax(T=T, L=lengths)
    if fused is None:
        fused = mucon_type == "flint" and flint_fusable(segmentation, lengths.shape[0])
    if fused:
        if mucon_type != "flint":
            raise ValueError("the fused kernel implements the flint loss only")
        E = flint_evidence(absolute_lengths, segmentation, [lengths.shape[0]], [T], overlap=overlap, template=template,
                           align_corners=align_corners)
        scaled = absolute_lengths * (1.0 + 2 * overlap)  # the in-place scaling of create_masks (masks.py:61)
        evidence = E / scaled[:, None]
        return F.nll_loss(F.log_softmax(evidence, dim=1), target_transcript, weight=class_weight, reduction="mean")
    masks = create_masks(T=T, L=absolute_lengths, template=template, overlap=overlap, align_corners=align_corners)
    return loss_from_masks(absolute_lengths, masks, segmentation, target_transcript, mucon_type, class_weight)


def mucon_loss_batch(lengths, segmentation, transcripts, Ms, Ts, template="box", overlap=0.0, class_weight=None,
                     align_corners=None, meta=None, mucon_type="flint"):
    """The mutual-consistency loss (models.py:414-525) of a packed batch, averaged over its videos: one
    fused-evidence launch for the whole batch instead of one Python loop per video and segment.
    lengths [sum Ms] s-head length logits (videos concatenated), segmentation [sum Ts, C] packed frame logits,
    transcripts [sum Ms] long.
    flint (:456-488), per video: F.nll_loss(log_softmax(E_v / L_v), transcript_v, reduction="mean"), E = masks @ seg.
    arithmetic (:489-523), per video: sum_i sum_t CE(seg_t, tr_i) * mask_i[t] / T = -sum_i (masks @ log_softmax(seg))[i,
    tr_i] / T -- the same evidence kernel applied to the log-probabilities, so the masks are not materialised either."""
    Ms_np, Ts_np = np.asarray(Ms, dtype=np.int64), np.asarray(Ts, dtype=np.int64)
    V, dev = int(Ms_np.shape[0]), segmentation.device
    meta = meta if meta is not None else _flint_meta(Ms_np, Ts_np, dev)
    if "row_vid64" not in meta:   # small index tables, built once per batch shape (no H2D copies in a captured step)
        meta["row_vid64"] = meta["row_vid"].long()
        meta["col"] = torch.from_numpy(np.concatenate([np.arange(m) for m in Ms_np]) if V else np.zeros(0, np.int64)).to(dev)
        meta["Tf"] = torch.from_numpy(Ts_np.astype(np.float32)).to(dev)
    row_vid, col, Tt = meta["row_vid64"], meta["col"], meta["Tf"]
    maxM = int(Ms_np.max(initial=1))
    # project_lengths_softmax per video (masks.py:8-12): softmax over the video's segments, times T
    padded = torch.full((V, maxM), float("-inf"), dtype=lengths.dtype, device=dev)
    padded = padded.index_put((row_vid, col), lengths)
    absolute = (F.softmax(padded, dim=1) * Tt[:, None])[row_vid, col]
    if mucon_type == "arithmetic":
        E = flint_evidence(absolute, F.log_softmax(segmentation, dim=1), Ms_np, Ts_np, overlap=overlap,
                           template=template, align_corners=align_corners, meta=meta)
        ce = -E.gather(1, transcripts.long()[:, None])[:, 0]
        if class_weight is not None:
            ce = ce * class_weight[transcripts.long()]
        per_video = torch.zeros(V, dtype=ce.dtype, device=dev).index_add_(0, row_vid, ce) / Tt
        return per_video.mean()
    if mucon_type != "flint":
        raise Exception(f"Invalid mucon type ({mucon_type})")
    E = flint_evidence(absolute, segmentation, Ms_np, Ts_np, overlap=overlap, template=template,
                       align_corners=align_corners, meta=meta)
    scaled = absolute * (1.0 + 2 * overlap)  # the in-place scaling of create_masks (masks.py:61)
    logp = F.log_softmax(E / scaled[:, None], dim=1)
    nll = -logp.gather(1, transcripts.long()[:, None])[:, 0]
    w = class_weight[transcripts.long()] if class_weight is not None else torch.ones_like(nll)
    num = torch.zeros(V, dtype=nll.dtype, device=dev).index_add_(0, row_vid, nll * w)
    den = torch.zeros(V, dtype=nll.dtype, device=dev).index_add_(0, row_vid, w)
    return (num / den).mean()


def smoothing_loss_packed(logits, Ts, log_softmax_before=True, clamp=True, clamp_min=0.0, clamp_max=16.0):
    """MuCon.calculate_smoothing_loss_for_logits (models.py:398-412) for a packed batch, averaged over its videos.
    Per video: values = F.mse_loss(x[1:], x[:-1].detach()) -- a scalar mean over (T-1) x C -- clamped (the reference
    clamps that scalar, not the elements), with x the (log-softmaxed) frame logits.  Frame pairs that straddle two
    videos are excluded."""
    Ts_np = np.asarray(Ts, dtype=np.int64)
    V, dev = int(Ts_np.shape[0]), logits.device
    x = F.log_softmax(logits, dim=1) if log_softmax_before else logits
    d = (x[1:] - x[:-1].detach()).pow(2).sum(1)                        # [sum T - 1]
    ends = np.cumsum(Ts_np)[:-1] - 1                                    # pair index (t, t+1) crossing a boundary
    keep = np.ones(max(int(Ts_np.sum()) - 1, 0), dtype=bool)
    keep[ends[ends >= 0]] = False
    vid = np.repeat(np.arange(V), Ts_np)[:-1] if Ts_np.sum() > 0 else np.zeros(0, np.int64)
    keep_t = torch.from_numpy(keep).to(dev)
    vid_t = torch.from_numpy(vid.astype(np.int64)).to(dev)
    per = torch.zeros(V, dtype=x.dtype, device=dev).index_add_(0, vid_t, d * keep_t)
    n = torch.from_numpy(np.maximum(Ts_np - 1, 1).astype(np.float32) * x.shape[1]).to(dev)
    per = per / n
    if clamp:
        per = torch.clamp(per, min=clamp_min, max=clamp_max)
    return per.mean()
