"""vit_mof -- ORACLE / TEST INFRASTRUCTURE ONLY.  Restates reference src/core/utils.py:34-47
(make_same_size_interpolate) and src/core/metrics/segmentation.py:16-44 (MoFAccuracyMetric)."""
import numpy as np
import torch
from torch.nn.functional import interpolate


def same_size_interpolate(prediction, t_len):
    p = torch.tensor([[np.asarray(prediction)]]).float()
    return interpolate(p, size=t_len, mode="nearest")[0][0].long().numpy()


def mof_counts(targets, predictions, ignore_ids=()):
    targets, predictions = np.asarray(targets), np.asarray(predictions)
    assert len(targets) == len(predictions)
    mask = np.logical_not(np.isin(targets, list(ignore_ids)))
    return int((targets[mask] == predictions[mask]).sum()), int(mask.sum())
