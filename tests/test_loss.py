"""Mutual-consistency loss (reference src/mucon/models.py:414-525): oracle vs the reference's frozen
outputs (CPU), CUDA-mask-based product function vs the same vectors (GPU).  Tolerance: rtol 2e-4 on the
loss and on grad(lengths) (float32 sums over T frames), 1e-3 relative on the frame-logit gradients."""
import os

import numpy as np
import pytest
import torch

from oracle import loss as oloss

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss.npz"))
CASES = [c.split(",") for c in G["cases"]]


def _case(i):
    T, N, C, mtype, tmpl, ov = CASES[i]
    return int(T), int(N), int(C), mtype, tmpl, float(ov)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_reference(i):
    T, N, C, mtype, tmpl, ov = _case(i)
    lengths = torch.from_numpy(G[f"c{i}_lengths"].copy()).requires_grad_(True)
    seg = torch.from_numpy(G[f"c{i}_seg"].copy()).requires_grad_(True)
    loss = oloss.mucon_loss(lengths, seg, torch.from_numpy(G[f"c{i}_tr"]), tmpl, ov, mtype)
    loss.backward()
    assert abs(loss.item() - float(G[f"c{i}_loss"])) <= 1e-5 * max(1.0, abs(float(G[f"c{i}_loss"])))
    assert np.allclose(lengths.grad.numpy(), G[f"c{i}_glen"], rtol=1e-4, atol=1e-6)
    assert np.allclose(seg.grad.numpy()[::37], G[f"c{i}_gseg_rows"], rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_cuda_loss_matches_reference(cuda_device, i):
    from mucon_b200.loss import mucon_loss
    T, N, C, mtype, tmpl, ov = _case(i)
    lengths = torch.from_numpy(G[f"c{i}_lengths"].copy()).to(cuda_device).requires_grad_(True)
    seg = torch.from_numpy(G[f"c{i}_seg"].copy()).to(cuda_device).requires_grad_(True)
    loss = mucon_loss(lengths, seg, torch.from_numpy(G[f"c{i}_tr"]).to(cuda_device), tmpl, ov, mtype)
    loss.backward()
    want = float(G[f"c{i}_loss"])
    assert abs(loss.item() - want) <= 2e-4 * max(1.0, abs(want))
    gl, wl = lengths.grad.cpu().numpy(), G[f"c{i}_glen"]
    assert np.allclose(gl, wl, rtol=2e-3, atol=2e-3 * np.abs(wl).max()), (gl, wl)
    gs = seg.grad.cpu().numpy()
    assert np.allclose(gs[::37], G[f"c{i}_gseg_rows"], rtol=1e-3, atol=1e-3 * np.abs(G[f"c{i}_gseg_rows"]).max())
    assert abs(np.abs(gs).astype(np.float64).sum() - float(G[f"c{i}_gseg_sum"])) <= 1e-3 * float(G[f"c{i}_gseg_sum"])
