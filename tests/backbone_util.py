"""Shared by the backbone tests: rebuilds the seeded weights / inputs of tests/golden/backbone.npz."""
import os

import numpy as np
import torch
import torch.nn as nn

STAGES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
POOL = [1, 2, 4, 8]
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backbone.npz"))
CASES = [tuple(int(x) for x in c.split(",")) for c in G["cases"]]


def build_model(D, H, C, seed):
    """Same construction order as tests/golden/make_golden_backbone.py -> identical seeded weights."""
    from mucon_b200.temporal import WaveNetBlock
    torch.manual_seed(seed)
    ft = WaveNetBlock(D, stages=STAGES, out_dims=H, pooling=True, pooling_type="max", pooling_layers=POOL,
                      leaky=False, dropout_rate=0.25).eval()
    gn = nn.GroupNorm(num_groups=32, num_channels=H).eval()
    cls = nn.Conv1d(H, C, kernel_size=1).eval()
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5)
        gn.bias.uniform_(-0.5, 0.5)
    return ft, gn, cls


def state_dict_of(ft, gn, cls):
    sd = {"ft." + k: v for k, v in ft.state_dict().items()}
    sd.update({"ft_last_gn." + k: v for k, v in gn.state_dict().items()})
    sd.update({"conv_classifier." + k: v for k, v in cls.state_dict().items()})
    return sd


def case_inputs(i):
    T, D, H, C, seed = CASES[i]
    ft, gn, cls = build_model(D, H, C, seed)
    g = torch.Generator().manual_seed(100 + seed)
    feats = torch.randn(1, T, D, generator=g).abs() * 0.5
    wsum = sum(p.double().abs().sum().item() for p in list(ft.parameters()) + list(gn.parameters()) + list(cls.parameters()))
    fresh = abs(wsum - float(G[f"c{i}_wsum"])) < 1e-6 * abs(wsum) and \
        abs(feats.double().sum().item() - float(G[f"c{i}_xsum"])) < 1e-6 * abs(float(G[f"c{i}_xsum"]))
    return (T, D, H, C), (ft, gn, cls), feats, fresh
