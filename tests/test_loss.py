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
@pytest.mark.parametrize("fused", [False, None])
@pytest.mark.parametrize("i", range(len(CASES)))
def test_cuda_loss_matches_reference(cuda_device, i, fused):
    """fused=False materialises the masks like the reference; None lets the flint cases go through the
    fused evidence kernel (mucon_flint_fwd / _bwd), which never writes the masks."""
    from mucon_b200.loss import flint_fusable, mucon_loss
    T, N, C, mtype, tmpl, ov = _case(i)
    lengths = torch.from_numpy(G[f"c{i}_lengths"].copy()).to(cuda_device).requires_grad_(True)
    seg = torch.from_numpy(G[f"c{i}_seg"].copy()).to(cuda_device).requires_grad_(True)
    if fused is None and not (mtype == "flint" and flint_fusable(seg, N)):
        pytest.skip("not a fused-kernel case")
    loss = mucon_loss(lengths, seg, torch.from_numpy(G[f"c{i}_tr"]).to(cuda_device), tmpl, ov, mtype, fused=fused)
    loss.backward()
    want = float(G[f"c{i}_loss"])
    assert abs(loss.item() - want) <= 2e-4 * max(1.0, abs(want))
    gl, wl = lengths.grad.cpu().numpy(), G[f"c{i}_glen"]
    assert np.allclose(gl, wl, rtol=2e-3, atol=2e-3 * np.abs(wl).max()), (gl, wl)
    gs = seg.grad.cpu().numpy()
    assert np.allclose(gs[::37], G[f"c{i}_gseg_rows"], rtol=1e-3, atol=1e-3 * np.abs(G[f"c{i}_gseg_rows"]).max())
    assert abs(np.abs(gs).astype(np.float64).sum() - float(G[f"c{i}_gseg_sum"])) <= 1e-3 * float(G[f"c{i}_gseg_sum"])


@pytest.mark.gpu
@pytest.mark.parametrize("tmpl,ov,align", [("box", 0.0, False), ("box", 0.15, False), ("gaussian", 0.1, False),
                                           ("trapezoid", 0.0, True), ("box", 0.0, True)])
def test_fused_evidence_equals_masks_times_logits(cuda_device, tmpl, ov, align):
    """flint_evidence on a ragged batch against (create_masks @ seg) per video: values and both gradients."""
    from mucon_b200.loss import flint_evidence
    from mucon_b200.masks import create_masks
    rng = np.random.default_rng(23)
    Ts = [2000, 317, 64, 5003, 9, 1000, 513, 512]
    Ms = [6, 3, 2, 12, 1, 30, 4, 5]
    Cn = 48
    Ls = [(rng.dirichlet(2 * np.ones(m)) * t * rng.uniform(0.9, 0.99)).astype(np.float32) for t, m in zip(Ts, Ms)]
    seg0 = rng.standard_normal((sum(Ts), Cn)).astype(np.float32)
    gE = rng.standard_normal((sum(Ms), Cn)).astype(np.float32)
    L = torch.from_numpy(np.concatenate(Ls)).to(cuda_device).requires_grad_(True)
    seg = torch.from_numpy(seg0).to(cuda_device).requires_grad_(True)
    E = flint_evidence(L, seg, Ms, Ts, overlap=ov, template=tmpl, align_corners=align)
    (E * torch.from_numpy(gE).to(cuda_device)).sum().backward()
    Lr = torch.from_numpy(np.concatenate(Ls)).to(cuda_device).requires_grad_(True)
    segr = torch.from_numpy(seg0).to(cuda_device).requires_grad_(True)
    outs, so, lo = [], 0, 0
    for t, m in zip(Ts, Ms):
        masks = create_masks(t, Lr[lo:lo + m] * 1.0, overlap=ov, template=tmpl, align_corners=align)
        outs.append(masks @ segr[so:so + t])
        so, lo = so + t, lo + m
    Er = torch.cat(outs)
    (Er * torch.from_numpy(gE).to(cuda_device)).sum().backward()
    scale = Er.abs().max().item()
    assert torch.allclose(E, Er, rtol=2e-4, atol=2e-4 * scale), (E - Er).abs().max().item()
    gs, gsr = seg.grad, segr.grad
    assert torch.allclose(gs, gsr, rtol=1e-4, atol=1e-4 * gsr.abs().max().item())
    gl, glr = L.grad.cpu().numpy(), Lr.grad.cpu().numpy()
    assert np.allclose(gl, glr, rtol=3e-3, atol=3e-3 * np.abs(glr).max()), np.abs(gl - glr).max()


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_batched_loss_matches_reference_single_videos(cuda_device, i):
    """mucon_loss_batch (flint and arithmetic through the fused evidence kernel, masks never written) on each frozen
    reference case as a batch of one"""
    from mucon_b200.loss import mucon_loss_batch
    T, N, C, mtype, tmpl, ov = _case(i)
    lengths = torch.from_numpy(G[f"c{i}_lengths"].copy()).to(cuda_device).requires_grad_(True)
    seg = torch.from_numpy(G[f"c{i}_seg"].copy()).to(cuda_device).requires_grad_(True)
    tr = torch.from_numpy(G[f"c{i}_tr"]).to(cuda_device)
    loss = mucon_loss_batch(lengths, seg, tr, [N], [T], template=tmpl, overlap=ov, mucon_type=mtype)
    loss.backward()
    want = float(G[f"c{i}_loss"])
    assert abs(loss.item() - want) <= 2e-4 * max(1.0, abs(want)), (loss.item(), want)
    gl, wl = lengths.grad.cpu().numpy(), G[f"c{i}_glen"]
    assert np.allclose(gl, wl, rtol=2e-3, atol=2e-3 * np.abs(wl).max()), (gl, wl)
    gs = seg.grad.cpu().numpy()
    assert np.allclose(gs[::37], G[f"c{i}_gseg_rows"], rtol=1e-3, atol=1e-3 * np.abs(G[f"c{i}_gseg_rows"]).max())


@pytest.mark.gpu
def test_batched_loss_is_mean_of_single_video_losses(cuda_device):
    from mucon_b200.loss import mucon_loss, mucon_loss_batch, smoothing_loss_packed
    rng = np.random.default_rng(3)
    Ts, Ms, C = [700, 333, 1200, 90], [6, 4, 9, 2], 24
    segs = [torch.from_numpy(rng.standard_normal((t, C)).astype(np.float32) * 2).to(cuda_device) for t in Ts]
    lens = [torch.from_numpy(rng.standard_normal(m).astype(np.float32)).to(cuda_device) for m in Ms]
    trs = [torch.from_numpy(rng.integers(0, C, m)).to(cuda_device) for m in Ms]
    for mtype in ("flint", "arithmetic"):
        want = sum(mucon_loss(l, s, t, "box", 0.0, mtype, fused=False) for l, s, t in zip(lens, segs, trs)) / len(Ts)
        got = mucon_loss_batch(torch.cat(lens), torch.cat(segs), torch.cat(trs), Ms, Ts, mucon_type=mtype)
        assert abs(got.item() - want.item()) <= 2e-4 * max(1.0, abs(want.item())), (mtype, got.item(), want.item())
    # smoothing loss (models.py:398-412): per video, the reference's statements
    import torch.nn.functional as F
    want = 0.0
    for s in segs:
        x = F.log_softmax(s, dim=1)
        want = want + torch.clamp(F.mse_loss(x[1:], x[:-1].detach()), min=0.0, max=16.0)
    got = smoothing_loss_packed(torch.cat(segs), Ts)
    assert abs(got.item() - (want / len(Ts)).item()) <= 1e-5 * max(1.0, abs(got.item()))
