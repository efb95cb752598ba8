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


# ---- segment-level metrics of the Viterbi head (evaluators.py:230-243) ---------------------------------------------
def runs(labels, ignore_ids=()):
    """Maximal runs of equal labels whose label is not in ignore_ids: (label, start, end) arrays.
    isba_code.py:11-21,36-44 (segment_labels / segment_intervals + the bg filter) and
    mstcn_code.py:6-27 (get_labels_start_end_time) produce exactly these."""
    y = np.asarray(labels)
    idx = np.concatenate([[0], np.nonzero(np.diff(y))[0] + 1, [len(y)]])
    lab, st, en = y[idx[:-1]], idx[:-1], idx[1:]
    keep = ~np.isin(lab, list(ignore_ids))
    return lab[keep], st[keep], en[keep]


def _overlap(P, Y, ignore_ids, union_of_both):
    """isba_code.py:22-61 (IoD) / :64-109 (IoU): per true segment the best score over predicted segments of the same
    label, score = intersection / (predicted length | span of both); mean over the true segments (NaN if none)."""
    tl, ts, te = runs(Y, ignore_ids)
    pl, ps, pe = runs(P, ignore_ids)
    scores = np.zeros(len(tl), dtype=np.float64)
    for i in range(len(tl)):
        for j in range(len(pl)):
            if tl[i] == pl[j]:
                inter = min(pe[j], te[i]) - max(ps[j], ts[i])
                den = (max(pe[j], te[i]) - min(ps[j], ts[i])) if union_of_both else (pe[j] - ps[j])
                scores[i] = max(scores[i], float(inter) / float(den))
    return scores.mean() if len(scores) else float("nan")


def iod(P, Y, ignore_ids=()):
    return _overlap(P, Y, ignore_ids, False)


def iou(P, Y, ignore_ids=()):
    return _overlap(P, Y, ignore_ids, True)


def levenshtein(p, y):
    """mstcn_code.py:30-50 (unnormalised distance)."""
    m, n = len(p), len(y)
    D = np.zeros((m + 1, n + 1), dtype=np.int64)
    D[:, 0] = np.arange(m + 1)
    D[0, :] = np.arange(n + 1)
    for j in range(1, n + 1):
        for i in range(1, m + 1):
            D[i, j] = D[i - 1, j - 1] if y[j - 1] == p[i - 1] else min(D[i - 1, j], D[i, j - 1], D[i - 1, j - 1]) + 1
    return int(D[m, n])


def edit_score(P, Y, ignore_ids=()):
    """mstcn_code.py:53-56 with norm=True: (1 - D / max(m, n)) * 100."""
    pl, _, _ = runs(P, ignore_ids)
    yl, _, _ = runs(Y, ignore_ids)
    d = levenshtein(pl, yl)
    return (1 - d / max(len(pl), len(yl))) * 100 if max(len(pl), len(yl)) else float("nan")


def f_counts(P, Y, overlap, ignore_ids=()):
    """mstcn_code.py:59-81: greedy matching of predicted to true segments by IoU -> (tp, fp, fn)."""
    pl, ps, pe = runs(P, ignore_ids)
    yl, ys, ye = runs(Y, ignore_ids)
    tp = fp = 0
    hits = np.zeros(len(yl), dtype=bool)
    for j in range(len(pl)):
        if len(yl) == 0:
            fp += 1
            continue
        inter = np.minimum(pe[j], ye) - np.maximum(ps[j], ys)
        union = np.maximum(pe[j], ye) - np.minimum(ps[j], ys)
        v = (1.0 * inter / union) * (pl[j] == yl)
        k = int(np.argmax(v))
        if v[k] >= overlap and not hits[k]:
            tp += 1
            hits[k] = True
        else:
            fp += 1
    return float(tp), float(fp), float(len(yl) - hits.sum())


def f1(tp, fp, fn):
    """fully_supervised.py:74-88."""
    prec, rec = (tp / (tp + fp), tp / (tp + fn)) if tp + fp != 0.0 else (0.0, 0.0)
    return 2.0 * prec * rec / (prec + rec) * 100 if prec + rec != 0.0 else 0.0
