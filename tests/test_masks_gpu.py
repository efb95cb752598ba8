"""GPU parity of mucon_masks_fwd / mucon_masks_bwd (through the reference's create_masks signature)
against the reference's frozen outputs and the torch oracle.

Tolerances: forward atol = 3e-5 * max(1, T / min L) (tests/util.mask_atol) -- the template
coordinate is computed in float32 from gx = g*s + x like the reference does, whose rounding error
scales with T/L; the kernel's and torch's op order differ in the last bit of gx.
grad_L: rtol 1e-3, atol 2e-3 * max|grad|.
"""
import os

import numpy as np
import pytest
import torch

from oracle import masks as omasks
from tests.util import mask_atol

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "masks.npz"))
CASES = [c.split(",") for c in G["cases"]]


@pytest.mark.parametrize("i", range(len(CASES)))
def test_forward_backward_match_reference_golden(cuda_device, i):
    from mucon_b200.masks import create_masks
    T, M, ov, tmpl = int(CASES[i][0]), int(CASES[i][1]), float(CASES[i][2]), CASES[i][3]
    g = {k: G[f"c{i}_{k}"] for k in ("L", "masks", "L_after", "gout", "gradL")}
    leaf = torch.from_numpy(g["L"].copy()).to(cuda_device).requires_grad_(True)
    Lw = leaf * 1.0
    masks = create_masks(T, Lw, overlap=ov, template=tmpl)
    assert masks.shape == (M, T)
    assert np.abs(masks.detach().cpu().numpy() - g["masks"]).max() <= mask_atol(T, g["L"])
    assert np.allclose(Lw.detach().cpu().numpy(), g["L_after"], rtol=1e-6)  # in-place scaling, masks.py:61
    (masks * torch.from_numpy(g["gout"]).to(cuda_device)).sum().backward()
    got = leaf.grad.cpu().numpy()
    scale = np.abs(g["gradL"]).max()
    assert np.allclose(got, g["gradL"], rtol=1e-3, atol=2e-3 * scale), (got, g["gradL"])


@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("tmpl", ["box", "gaussian", "trapezoid"])
def test_both_sampling_conventions_against_torch_oracle(cuda_device, align, tmpl):
    from mucon_b200.masks import create_masks
    rng = np.random.default_rng(3)
    for T, M, ov in [(2000, 6, 0.0), (333, 5, 0.2), (4096, 12, 0.05)]:
        L0 = (rng.dirichlet(2 * np.ones(M)) * T).astype(np.float32)
        ref_L = torch.from_numpy(L0.copy()).requires_grad_(True)
        ref, _ = omasks.create_masks_torch(T, ref_L, ov, tmpl, align_corners=align)
        gout = torch.from_numpy(rng.standard_normal((M, T)).astype(np.float32))
        (ref * gout).sum().backward()
        leaf = torch.from_numpy(L0.copy()).to(cuda_device).requires_grad_(True)
        out = create_masks(T, leaf * 1.0, overlap=ov, template=tmpl, align_corners=align)
        assert np.abs(out.detach().cpu().numpy() - ref.detach().numpy()).max() <= mask_atol(T, L0 * (1 + 2 * ov))
        (out * gout.to(cuda_device)).sum().backward()
        scale = ref_L.grad.abs().max().item()
        assert np.allclose(leaf.grad.cpu().numpy(), ref_L.grad.numpy(), rtol=1e-3, atol=2e-3 * scale)


@pytest.mark.parametrize("align", [False, True])
def test_box_ranges_on_extreme_rows(cuda_device, align):
    """The box template is written through certified constant ranges (csrc/masks.cu make_regions):
    rows with windows shorter than a frame, longer than the video, hanging over either end, tiny
    and large T, with and without overlap -- forward and gradient against the torch oracle."""
    from mucon_b200.masks import create_masks
    rng = np.random.default_rng(17)
    cases = []
    for T in (1, 2, 7, 8, 9, 33, 100, 257, 1000, 5003):
        if T == 1 and align:
            continue  # unit-size grids with align_corners=True changed meaning in torch 1.3 (the oracle warns)
        for _ in range(6):
            M = int(rng.integers(1, 9))
            kind = rng.integers(0, 4)
            if kind == 0:
                L0 = rng.dirichlet(0.3 * np.ones(M)) * T            # some segments far below one frame
            elif kind == 1:
                L0 = rng.uniform(0.05, 3.0, M) * T                    # windows longer than the video
            elif kind == 2:
                L0 = np.full(M, T / M)
            else:
                L0 = rng.uniform(0.01, 1.0, M) * T / M
            # keep window edges off the frame grid: with sum(L) == T the last frame sits exactly on the
            # last window's edge (align_corners=True), where the gradient is discontinuous and the last
            # bit of u decides it
            L0 = np.maximum(L0 * rng.uniform(0.9, 0.99), 1e-2).astype(np.float32)
            cases.append((T, M, float(rng.choice([0.0, 0.1, 0.5])), L0))
    for T, M, ov, L0 in cases:
        ref_L = torch.from_numpy(L0.copy()).requires_grad_(True)
        ref, _ = omasks.create_masks_torch(T, ref_L, ov, "box", align_corners=align)
        gout = torch.from_numpy(rng.standard_normal((M, T)).astype(np.float32))
        (ref * gout).sum().backward()
        leaf = torch.from_numpy(L0.copy()).to(cuda_device).requires_grad_(True)
        out = create_masks(T, leaf * 1.0, overlap=ov, template="box", align_corners=align)
        err = np.abs(out.detach().cpu().numpy() - ref.detach().numpy()).max()
        assert err <= mask_atol(T, L0 * (1 + 2 * ov)), (T, M, ov, L0, err)
        (out * gout.to(cuda_device)).sum().backward()
        scale = max(ref_L.grad.abs().max().item(), 1e-6)
        # + 1e-2: torch's own backward leaves cancellation noise of that order where the true gradient is 0
        assert np.allclose(leaf.grad.cpu().numpy(), ref_L.grad.numpy(), rtol=2e-3, atol=4e-3 * scale + 1e-2), (T, M, ov, L0)


def test_batched_launch_equals_per_video(cuda_device):
    from mucon_b200.masks import create_masks, create_masks_batch
    rng = np.random.default_rng(5)
    Ts = [2000, 317, 1, 9000, 64]
    Ms = [6, 3, 2, 12, 1]
    Ls = [(rng.dirichlet(np.ones(m)) * t).astype(np.float32) for t, m in zip(Ts, Ms)]
    flat = torch.from_numpy(np.concatenate(Ls)).to(cuda_device)
    out, off = create_masks_batch(Ts, flat, Ms)
    for v in range(len(Ts)):
        one = create_masks(Ts[v], torch.from_numpy(Ls[v].copy()).to(cuda_device))
        assert torch.equal(out[off[v]:off[v + 1]].view(Ms[v], Ts[v]), one)


def test_bad_inputs(cuda_device):
    from mucon_b200 import _lib
    from mucon_b200.masks import create_masks
    with pytest.raises(NameError):
        create_masks(10, torch.ones(2, device=cuda_device), template="triangle")
    with pytest.raises(_lib.MuconError):
        create_masks(10, torch.ones(2))
