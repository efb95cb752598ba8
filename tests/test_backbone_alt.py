"""Alternate backbones selected by model.ft.type (reference src/core/modules/temporal.py:150-204 MS-TCN++
first stage, :56-74 NoFt).  CPU: the host-side weight folding is exact linear algebra.  GPU: the
tcgen05 path against the frozen outputs of the unmodified reference modules
(tests/golden/make_golden_backbone_alt.py), atol = 1e-2 * RMS of the reference tensor (TF32 operands)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "backbone_alt.npz"))
CASES = [c.split(",") for c in G["cases"]]


def build(kind, L, D, seed):
    from mucon_b200.temporal import MSTCNPPFirstStage, NoFt
    torch.manual_seed(seed)
    if kind == "mstcnpp":
        return MSTCNPPFirstStage(num_layers=L, num_f_maps=128, input_dim=D, output_dim=128).eval()
    return NoFt(in_chnnels=D, out_dims=128).eval()


def reference_style_forward(m, x):
    """temporal.py:182-203 with the module's own parameters, float64, eval mode."""
    f = F.conv1d(x, m.conv_1x1_in.weight.double(), m.conv_1x1_in.bias.double())
    for i in range(m.num_layers):
        d1, d2 = 2 ** (m.num_layers - 1 - i), 2 ** i
        a = F.conv1d(f, m.conv_dilated_1[i].weight.double(), m.conv_dilated_1[i].bias.double(), padding=d1, dilation=d1)
        b = F.conv1d(f, m.conv_dilated_2[i].weight.double(), m.conv_dilated_2[i].bias.double(), padding=d2, dilation=d2)
        g = F.conv1d(torch.cat([a, b], 1), m.conv_fusion[i].weight.double(), m.conv_fusion[i].bias.double())
        f = F.relu(g) + f
        if i in m.pooling_layers:
            f = F.max_pool1d(f, kernel_size=2)
    return F.conv1d(f, m.conv_out.weight.double(), m.conv_out.bias.double())


def folded_forward(m, x):
    """The same network through the folded per-layer weights the GPU path uses (float64 on the CPU)."""
    w = m._weights()
    f = F.conv1d(x, m.conv_1x1_in.weight.double(), m.conv_1x1_in.bias.double())
    for i, (shifts, W, bias) in enumerate(w["layers"]):
        T = f.shape[2]
        acc = bias.double()[None, :, None].expand(1, 128, T).clone()
        for j, sft in enumerate(shifts):
            Ws = W[j * 128:(j + 1) * 128].double()          # [Co, Ci]
            src = torch.zeros_like(f)
            lo, hi = max(0, -sft), min(T, T - sft)
            if hi > lo:
                src[:, :, lo:hi] = f[:, :, lo + sft:hi + sft]
            acc = acc + torch.einsum("oc,bct->bot", Ws, src)
        f = F.relu(acc) + f
        if i in m.pooling_layers:
            f = F.max_pool1d(f, kernel_size=2)
    return F.conv1d(f, m.conv_out.weight.double(), m.conv_out.bias.double())


@pytest.mark.parametrize("L,T", [(11, 300), (6, 97), (1, 40), (4, 33)])
def test_weight_folding_is_exact_linear_algebra(L, T):
    m = build("mstcnpp", L, 32, seed=L)
    x = torch.randn(1, 32, T, dtype=torch.float64)
    with torch.no_grad():
        ref, got = reference_style_forward(m, x), folded_forward(m, x)
    assert ref.shape == got.shape
    assert (ref - got).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())  # folded weights are stored fp32
    shifts = [s for s, _, _ in m._weights()["layers"]]
    assert all(0 in s and len(s) <= 5 for s in shifts)


def test_constructor_and_state_dict_names_match_reference():
    ref_path = "/root/reference/src"
    if not os.path.isdir(ref_path):
        pytest.skip("reference tree not present")
    import sys
    sys.path.insert(0, ref_path)
    from core.modules.temporal import MSTCNPPFirstStage as RefStage, NoFt as RefNoFt
    mine, ref = build("mstcnpp", 5, 64, 0), RefStage(num_layers=5, num_f_maps=128, input_dim=64, output_dim=128)
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    assert list(build("noft", 0, 64, 0).state_dict().keys()) == list(RefNoFt(64, 128).state_dict().keys())


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_alternate_backbones_against_reference_outputs(cuda_device, i):
    kind, L, D, T, seed = CASES[i][0], int(CASES[i][1]), int(CASES[i][2]), int(CASES[i][3]), int(CASES[i][4])
    m = build(kind, L, D, seed)
    wsum = sum(p.double().abs().sum().item() for p in m.parameters())
    assert wsum == pytest.approx(float(G[f"c{i}_wsum"]), rel=1e-12), "seeded weights drifted (torch RNG changed?)"
    g = torch.Generator().manual_seed(200 + seed)
    feats = torch.randn(1, T, D, generator=g).abs() * 0.5
    assert feats.double().sum().item() == pytest.approx(float(G[f"c{i}_xsum"]), rel=1e-12)
    z = m.to(cuda_device)(feats.permute(0, 2, 1).to(cuda_device))[0].t().cpu().numpy()
    ref = G[f"c{i}_z"]
    assert z.shape == ref.shape
    rms = float(np.sqrt(np.mean(np.square(ref))))
    assert np.abs(z - ref).max() <= 1e-2 * rms, (np.abs(z - ref).max(), rms)


@pytest.mark.gpu
def test_mstcnpp_ragged_batch_and_model_wrapper(cuda_device):
    """A ragged packed batch (tile-boundary lengths) equals the videos run one by one, and
    MuConBackbone(ft_type=...) wires the alternates like models.py:160-186."""
    from mucon_b200.temporal import MuConBackbone
    torch.manual_seed(5)
    net = MuConBackbone(input_feature_size=64, num_classes=20, ft_type="mstcnpp").eval().to(cuda_device)
    Ts = [700, 16, 129, 128, 127, 1999, 33]
    feats = [torch.randn(t, 64).abs().to(cuda_device) for t in Ts]
    plan = net.plan(Ts)
    z = net.encode_packed(torch.cat(feats), plan)
    lp = net.logprobs_packed(z, plan)
    assert lp.shape == (sum(Ts), 20)
    off = plan.off_host[-1]
    for v, t in enumerate(Ts):
        p1 = net.plan([t])
        z1 = net.encode_packed(feats[v], p1)
        assert torch.allclose(z[off[v]:off[v + 1]], z1, rtol=1e-5, atol=1e-5)
    with pytest.raises(Exception):
        MuConBackbone(ft_type="nope")
    nf = MuConBackbone(input_feature_size=64, num_classes=20, ft_type="noft").eval().to(cuda_device)
    assert nf.logprobs_packed(nf.encode_packed(torch.cat(feats), nf.plan(Ts)), nf.plan(Ts)).shape == (sum(Ts), 20)
