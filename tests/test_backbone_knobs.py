"""model.ft.leaky_relu and model.ft.pooling_type = "avg" (reference src/configs/mucon/default.py:81-96,
src/core/modules/temporal.py:35-41,98-101,139-142) against outputs of the unmodified reference WaveNetBlock
(tests/golden/backbone_knobs.npz).  TF32 tensor-core path: max |err| <= 2e-2 * RMS of the reference tensor; with
tensor_cores=False the 128 -> 128 convolutions run on the exact fp32 kernels and only the input projection reads
TF32 operands: max |err| <= 5e-3 * RMS (measured 6e-4 absolute)."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backbone_knobs.npz"))
STAGES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
POOL = [1, 2, 4, 8]
CASES = [c.split(",") for c in G["cases"]]


def build(i):
    from mucon_b200.temporal import WaveNetBlock
    T, D, leaky, ptype, seed = int(CASES[i][0]), int(CASES[i][1]), CASES[i][2] == "True", CASES[i][3], int(CASES[i][4])
    torch.manual_seed(seed)
    ft = WaveNetBlock(D, stages=STAGES, out_dims=128, pooling=True, pooling_type=ptype, pooling_layers=POOL, leaky=leaky,
                      dropout_rate=0.25).eval()
    g = torch.Generator().manual_seed(300 + seed)
    feats = torch.randn(1, T, D, generator=g).abs() * 0.5
    wsum = sum(p.double().abs().sum().item() for p in ft.parameters())
    fresh = abs(wsum - float(G[f"c{i}_wsum"])) < 1e-6 * wsum and \
        abs(feats.double().sum().item() - float(G[f"c{i}_xsum"])) < 1e-6 * abs(float(G[f"c{i}_xsum"]))
    return ft, feats, fresh, (leaky, ptype)


def test_knob_constructor_accepts_reference_configuration():
    ft, _, _, (leaky, ptype) = build(1)
    assert ft.leaky == leaky and ft.pooling_type == ptype and len(list(ft.parameters())) == 2 * (2 * 11 + 2)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
@pytest.mark.parametrize("tensor_cores", [False, True])
def test_knobs_match_reference(cuda_device, i, tensor_cores):
    from mucon_b200.temporal import BackbonePlan
    ft, feats, fresh, _ = build(i)
    if not fresh:
        pytest.skip("seeded weights / inputs differ from the ones the golden file was minted with")
    ft = ft.to(cuda_device)
    T = feats.shape[1]
    plan = BackbonePlan([T], ft.n_pools(), cuda_device)
    z = ft.forward_packed(feats[0].to(cuda_device).contiguous(), plan, tensor_cores=tensor_cores).cpu().numpy()
    want = G[f"c{i}_z"]
    assert z.shape == want.shape
    rms = float(np.sqrt(np.mean(want.astype(np.float64) ** 2)))
    assert np.abs(z - want).max() <= (2e-2 if tensor_cores else 5e-3) * rms, (np.abs(z - want).max(), rms)
