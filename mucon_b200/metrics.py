"""vit_mof on the device: the step right after the Viterbi decode in the reference evaluator
(src/mucon/evaluators.py:225-243): make_same_size_interpolate (src/core/utils.py:34-47) +
MoFAccuracyMetric (src/core/metrics/segmentation.py:16-44).  Only per-video counters leave the GPU."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def mof_counts(pred, pred_off, gt, gt_off, ignore_ids=()):
    """pred: int32 CUDA tensor of predicted frame labels (videos concatenated, offsets pred_off [V+1]);
    gt: int32 CUDA tensor of targets with offsets gt_off.  Returns an int64 tensor [V, 2] (correct, total)."""
    if not pred.is_cuda or not gt.is_cuda:
        raise _lib.MuconError("mof_counts needs CUDA tensors (there is no CPU fallback)")
    pred_off = torch.as_tensor(np.asarray(pred_off, dtype=np.int64))
    gt_off_h = np.asarray(gt_off, dtype=np.int64)
    V = int(pred_off.shape[0]) - 1
    dev = pred.device
    po, go = pred_off.to(dev), torch.from_numpy(gt_off_h).to(dev)
    counts = torch.empty((V, 2), dtype=torch.int64, device=dev)
    ign = np.asarray(list(ignore_ids), dtype=np.int32)
    max_t = int(np.diff(gt_off_h).max(initial=0))
    _lib.check(_lib.lib().mucon_vit_mof(
        _lib.ptr(pred), _lib.ptr(po), _lib.ptr(gt), _lib.ptr(go), C.c_int(V), C.c_int(max_t),
        ign.ctypes.data_as(C.c_void_p) if ign.size else None, C.c_int(int(ign.size)), _lib.ptr(counts),
        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "mucon_vit_mof")
    return counts


def mof(counts):
    """MoFAccuracyMetric.summary(): sum(correct) / sum(total), 0.0 when nothing was counted."""
    c = counts.sum(0)
    return float(c[0]) / float(c[1]) if int(c[1]) else 0.0


def segment_metrics(pred, pred_off, gt, gt_off, ignore_ids=()):
    """IoD / IoU / Edit / F1-matching counts of every video on the device (isba_code.py:22-109, mstcn_code.py:27-81
    after make_same_size_interpolate; evaluators.py:225-243).  Same arguments as mof_counts.  Returns a float64 CUDA
    tensor [V, 12]: iod, iou, edit, then (tp, fp, fn) at overlaps 0.1, 0.25, 0.5."""
    if not pred.is_cuda or not gt.is_cuda:
        raise _lib.MuconError("segment_metrics needs CUDA tensors (there is no CPU fallback)")
    po_h = np.asarray(pred_off, dtype=np.int64)
    go_h = np.asarray(gt_off, dtype=np.int64)
    V = int(po_h.shape[0]) - 1
    dev = pred.device
    po, go = torch.from_numpy(po_h).to(dev), torch.from_numpy(go_h).to(dev)
    lib = _lib.lib()
    words = int(lib.mucon_vit_segment_metrics_ws_words(C.c_int64(int(go_h[-1])), C.c_int(V)))
    ws = torch.empty((words + 1) // 2, dtype=torch.int64, device=dev)  # 8-byte aligned
    out = torch.empty((V, 12), dtype=torch.float64, device=dev)
    ign = np.asarray(list(ignore_ids), dtype=np.int32)
    _lib.check(lib.mucon_vit_segment_metrics(
        _lib.ptr(pred), _lib.ptr(po), _lib.ptr(gt), _lib.ptr(go), C.c_int(V),
        ign.ctypes.data_as(C.c_void_p) if ign.size else None, C.c_int(int(ign.size)), _lib.ptr(ws), _lib.ptr(out),
        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "mucon_vit_segment_metrics")
    return out


def summarize(seg):
    """The evaluator's summaries of a [V, 12] segment_metrics tensor: IoDMetric / IoUMetric / Edit .summary() are means
    over videos (segmentation.py:78-82, fully_supervised.py:29-33), F1Score.summary() pools the counts
    (fully_supervised.py:64-88)."""
    s = seg.detach().cpu().numpy()
    if s.shape[0] == 0:
        return {"iod": 0.0, "iou": 0.0, "edit": 0.0, "f1": [0.0, 0.0, 0.0]}
    f1 = []
    for k in range(3):
        tp, fp, fn = (float(s[:, 3 + 3 * k + i].sum()) for i in range(3))
        prec, rec = (tp / (tp + fp), tp / (tp + fn)) if tp + fp != 0.0 else (0.0, 0.0)
        f1.append(2.0 * prec * rec / (prec + rec) * 100 if prec + rec != 0.0 else 0.0)
    return {"iod": float(sum(s[:, 0].tolist()) / s.shape[0]), "iou": float(sum(s[:, 1].tolist()) / s.shape[0]),
            "edit": float(np.array(s[:, 2]).mean()), "f1": f1}
