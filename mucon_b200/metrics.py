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
