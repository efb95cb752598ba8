"""model.ft.hidden_size = 64 / 32 (reference src/configs/mucon/default.py:87) on the 128-channel tensor-core kernels
through a zero-padded twin (WaveNetBlock._padded_twin): against outputs of the unmodified reference WaveNetBlock
(tests/golden/backbone_widths.npz, make_golden_backbone_widths.py).  Bars as for the 128-channel block
(tests/test_backbone_bf16.py): fp16 / tf32 max |err| <= 3e-2 / 2e-2 * RMS, RMS error <= 2e-3 * RMS."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backbone_widths.npz"))
STAGES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
POOL = [1, 2, 4, 8]
CASES = [[int(x) for x in c.split(",")] for c in G["cases"]]


def build(i):
    from mucon_b200.temporal import WaveNetBlock
    T, D, H, seed = CASES[i]
    torch.manual_seed(seed)
    ft = WaveNetBlock(D, stages=STAGES, out_dims=H, pooling=True, pooling_layers=POOL, dropout_rate=0.25).eval()
    g = torch.Generator().manual_seed(400 + seed)
    feats = torch.randn(1, T, D, generator=g).abs() * 0.5
    wsum = sum(p.detach().double().abs().sum().item() for p in ft.parameters())
    fresh = abs(wsum - float(G[f"c{i}_wsum"])) < 1e-6 * wsum and \
        abs(feats.double().sum().item() - float(G[f"c{i}_xsum"])) < 1e-6 * abs(float(G[f"c{i}_xsum"]))
    return ft, feats, fresh


def test_padded_twin_holds_the_parameters_and_stays_out_of_the_state_dict():
    ft, _, _ = build(0)
    keys = set(ft.state_dict().keys())
    tw = ft._padded_twin()
    assert set(ft.state_dict().keys()) == keys and tw.out_dims == 128 and ft._padded_twin() is tw
    h = ft.out_dims
    for a, b in zip(ft.layers + [ft], tw.layers + [tw]):
        for name in (("last_conv",) if a is ft else ("dilated_conv", "conv_1x1")):
            wa, wb = getattr(a, name).weight.detach(), getattr(b, name).weight.detach()
            assert torch.equal(wb[:h, :h], wa) and float(wb[h:].abs().max()) == 0 and float(wb[:, h:].abs().max()) == 0
            assert float(getattr(b, name).bias.detach()[h:].abs().max()) == 0
    assert torch.equal(tw.first_conv.weight.detach()[:h], ft.first_conv.weight.detach())
    with torch.no_grad():
        ft.first_conv.bias.add_(1.0)          # a parameter update invalidates the cached twin
    assert ft._padded_twin() is not tw


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
@pytest.mark.parametrize("precision", ["fp16", "tf32", "fp32"])
def test_narrow_blocks_match_reference(cuda_device, i, precision):
    from mucon_b200.temporal import BackbonePlan
    ft, feats, fresh = build(i)
    if not fresh:
        pytest.skip("seeded weights / inputs differ from the ones the golden file was minted with")
    ft = ft.to(cuda_device)
    T = feats.shape[1]
    plan = BackbonePlan([T], ft.n_pools(), cuda_device)
    z = ft.forward_packed(feats[0].to(cuda_device).contiguous(), plan, precision=precision).cpu().numpy()
    want = G[f"c{i}_z"]
    assert z.shape == want.shape
    rms = float(np.sqrt(np.mean(want.astype(np.float64) ** 2)))
    err = np.abs(z - want)
    print(f"\n[{precision}] hidden {ft.out_dims} T={T}: max|err|/RMS={err.max() / rms:.4f} rms(err)/RMS={np.sqrt(np.mean(err ** 2)) / rms:.5f}")
    max_bar = {"fp16": 3e-2, "tf32": 2e-2, "fp32": 5e-3}[precision]
    assert err.max() <= max_bar * rms and np.sqrt(np.mean(err ** 2)) <= 2e-3 * rms


@pytest.mark.gpu
def test_narrow_backbone_end_to_end_equals_its_fp32_path(cuda_device):
    """MuConBackbone(hidden_size=64): tensor-core twin + generic GroupNorm / classifier kernels against the same model
    on the exact fp32 kernels"""
    from mucon_b200.temporal import MuConBackbone
    torch.manual_seed(3)
    net = MuConBackbone(input_feature_size=256, num_classes=20, hidden_size=64).eval().to(cuda_device)
    Ts = [700, 333, 129, 2000]
    plan = net.plan(Ts, cuda_device)
    feats = (torch.randn(sum(Ts), 256, generator=torch.Generator().manual_seed(4)).abs() * 0.5).to(cuda_device)
    table, off = net.infer_pooled_packed(feats, plan)
    z32 = net.encode_packed(feats, plan, tensor_cores=False)
    want, _ = net.logprobs_pooled_packed(z32, plan)
    assert table.shape == want.shape and table.shape[1] == 20
    rms = float(want.pow(2).mean().sqrt())
    assert float((table - want).abs().max()) <= 3e-2 * rms
    assert float((table.argmax(1) == want.argmax(1)).float().mean()) >= 0.98
