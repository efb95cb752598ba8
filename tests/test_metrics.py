"""vit_mof counters: oracle vs the reference metric classes (CPU, build container), CUDA vs oracle (GPU)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import metrics as om

REF = "/root/reference/src"


def _cases(seed=0, n=12):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        Tp, Tg = int(rng.integers(30, 3000)), int(rng.integers(30, 3000))
        nseg = int(rng.integers(1, 9))
        cuts = np.sort(rng.choice(np.arange(1, Tp), nseg - 1, replace=False)) if nseg > 1 else np.array([], int)
        pred = np.repeat(rng.integers(0, 10, nseg), np.diff(np.concatenate([[0], cuts, [Tp]]))).astype(np.int32)
        gt = np.repeat(rng.integers(0, 10, 7), rng.multinomial(Tg, np.ones(7) / 7)).astype(np.int32)
        out.append((pred, gt))
    out.append((np.arange(5, dtype=np.int32), np.arange(5, dtype=np.int32)))  # same size
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_oracle_matches_reference_metric_classes():
    # the reference module needs the un-vendored fandak package only for tensor_to_numpy
    fandak = types.ModuleType("fandak"); futils = types.ModuleType("fandak.utils"); ftorch = types.ModuleType("fandak.utils.torch")
    ftorch.tensor_to_numpy = lambda t: t.detach().cpu().numpy()
    sys.modules.setdefault("fandak", fandak); sys.modules.setdefault("fandak.utils", futils)
    sys.modules.setdefault("fandak.utils.torch", ftorch)
    sys.path.insert(0, REF)
    try:
        from core.metrics.segmentation import MoFAccuracyMetric
        from core.utils import make_same_size_interpolate
        for ignore in ((), (0,), (0, 3)):
            m = MoFAccuracyMetric(ignore_ids=ignore)
            c_sum = t_sum = 0
            for pred, gt in _cases():
                same = make_same_size_interpolate(prediction=pred, target=gt)
                assert np.array_equal(same, om.same_size_interpolate(pred, len(gt)))
                m.add(targets=gt, predictions=same)
                c, t = om.mof_counts(gt, same, ignore)
                c_sum += c; t_sum += t
            assert m.correct == c_sum and m.total == t_sum
            assert m.summary() == (c_sum / t_sum if t_sum else 0.0)
    finally:
        sys.path.remove(REF)


@pytest.mark.gpu
def test_cuda_counters_match_oracle(cuda_device):
    from mucon_b200.metrics import mof, mof_counts
    cases = _cases(seed=3, n=40)
    pred = torch.from_numpy(np.concatenate([p for p, _ in cases])).to(cuda_device)
    gt = torch.from_numpy(np.concatenate([g for _, g in cases])).to(cuda_device)
    po = np.concatenate([[0], np.cumsum([len(p) for p, _ in cases])])
    go = np.concatenate([[0], np.cumsum([len(g) for _, g in cases])])
    for ignore in ((), (0,), (0, 3, 9)):
        counts = mof_counts(pred, po, gt, go, ignore).cpu().numpy()
        want = np.array([om.mof_counts(g, om.same_size_interpolate(p, len(g)), ignore) for p, g in cases])
        assert np.array_equal(counts, want)
        assert mof(torch.from_numpy(counts)) == (want[:, 0].sum() / want[:, 1].sum())
