"""Mints tests/golden/backbone.npz: outputs of the UNMODIFIED reference WaveNetBlock
(/root/reference/src/core/modules/temporal.py) followed by the GroupNorm / ReLU / nearest-interpolate
/ 1x1 classifier / log_softmax steps of src/mucon/models.py:746-773,567-582,368 (written out with the
same torch calls because mucon.models itself needs the un-vendored fandak package).

Weights and inputs are regenerated from seeds (torch.manual_seed) rather than stored; the file keeps
checksums of both so that a drifting RNG is detected instead of mis-reported as a parity failure."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))
from core.modules.temporal import WaveNetBlock  # noqa: E402

STAGES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
POOL = [1, 2, 4, 8]


def build(D, H, C, seed):
    torch.manual_seed(seed)
    ft = WaveNetBlock(in_channels=D, stages=STAGES, out_dims=H, pooling=True, pooling_type="max",
                      pooling_layers=POOL, leaky=False, dropout_rate=0.25).eval()
    gn = nn.GroupNorm(num_groups=32, num_channels=H).eval()
    cls = nn.Conv1d(H, C, kernel_size=1).eval()
    with torch.no_grad():  # non-trivial affine parameters
        gn.weight.uniform_(0.5, 1.5)
        gn.bias.uniform_(-0.5, 0.5)
    return ft, gn, cls


def main():
    out = {"torch_version": torch.__version__}
    cases = [(2000, 2048, 128, 48, 0), (777, 256, 128, 48, 1), (333, 64, 128, 20, 2), (125, 2048, 128, 48, 3)]
    for i, (T, D, H, C, seed) in enumerate(cases):
        ft, gn, cls = build(D, H, C, seed)
        g = torch.Generator().manual_seed(100 + seed)
        feats = torch.randn(1, T, D, generator=g).abs() * 0.5  # post-ReLU-like I3D features
        with torch.no_grad():
            z = ft(feats.permute(0, 2, 1))                 # models.py:753-756
            z = F.relu(gn(z))                               # models.py:759-764
            enc = z.permute(0, 2, 1)                        # [1, Tz, H]
            seg = cls(F.interpolate(z, T))                  # models.py:574-580
            logp = F.log_softmax(seg.squeeze(0).permute(1, 0), dim=1)   # models.py:346-348,368
        out[f"c{i}_z"] = enc[0].numpy()
        out[f"c{i}_logp"] = logp.numpy().astype(np.float32) if T <= 800 else logp.numpy()[::7].copy()
        out[f"c{i}_wsum"] = np.float64(sum(p.double().abs().sum().item() for p in list(ft.parameters()) + list(gn.parameters()) + list(cls.parameters())))
        out[f"c{i}_xsum"] = np.float64(feats.double().sum().item())
    out["cases"] = np.array([",".join(map(str, c)) for c in cases])
    np.savez_compressed(os.path.join(HERE, "backbone.npz"), **out)
    print("wrote backbone.npz")


if __name__ == "__main__":
    main()
