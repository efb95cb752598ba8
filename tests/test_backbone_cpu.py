"""CPU: the torch fp32 backbone oracle against the reference's frozen outputs (tests/golden/backbone.npz,
minted by tests/golden/make_golden_backbone.py from the unmodified reference WaveNetBlock)."""
import numpy as np
import pytest
import torch

from oracle import backbone as ob
from tests.backbone_util import CASES, G, POOL, STAGES, case_inputs, state_dict_of


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_reference_golden(i):
    (T, D, H, C), (ft, gn, cls), feats, fresh = case_inputs(i)
    if not fresh:
        pytest.skip("torch RNG stream differs from the one the fixture was minted with")
    sd = state_dict_of(ft, gn, cls)
    with torch.no_grad():
        z = ob.encode(sd, feats, STAGES, POOL)
        logp = ob.logprobs(sd, z, T)
    assert np.abs(z[0].numpy() - G[f"c{i}_z"]).max() <= 1e-5
    want = G[f"c{i}_logp"]
    got = logp.numpy() if T <= 800 else logp.numpy()[::7]
    assert np.abs(got - want).max() <= 1e-5


def test_state_dict_names_match_the_reference():
    from mucon_b200.temporal import MuConBackbone
    m = MuConBackbone(input_feature_size=64, num_classes=48)
    keys = set(m.state_dict().keys())
    for k in ("ft.first_conv.weight", "ft.l_0.dilated_conv.weight", "ft.l_10.conv_1x1.bias", "ft.last_conv.weight",
              "ft_last_gn.weight", "conv_classifier.weight"):
        assert k in keys
    assert sum(p.numel() for p in MuConBackbone().ft.parameters()) == 1002496


def test_nearest_index_matches_torch_interpolate():
    """The expansion index used by mucon_logsoftmax_expand: min(floor(t * (float)Tz / T), Tz - 1)."""
    import torch.nn.functional as F
    for T, Tz in [(2000, 125), (777, 48), (333, 20), (10000, 625), (301, 18), (125, 7), (9999, 624)]:
        src = torch.arange(Tz, dtype=torch.float32).view(1, 1, Tz)
        ref = F.interpolate(src, T)[0, 0].long().numpy()
        scale = np.float32(Tz) / np.float32(T)
        mine = np.minimum(np.floor(np.arange(T, dtype=np.float32) * scale).astype(np.int64), Tz - 1)
        assert np.array_equal(ref, mine), (T, Tz)
