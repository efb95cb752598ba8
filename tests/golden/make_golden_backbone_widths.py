"""Mints tests/golden/backbone_widths.npz: outputs of the UNMODIFIED reference WaveNetBlock
(/root/reference/src/core/modules/temporal.py:77-147) with model.ft.hidden_size = 64 and 32
(src/configs/mucon/default.py:87).  Weights / inputs are regenerated from seeds (checksums stored), as in
make_golden_backbone.py."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))
from core.modules.temporal import WaveNetBlock  # noqa: E402

STAGES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
POOL = [1, 2, 4, 8]
CASES = [(700, 256, 64, 0), (333, 2048, 64, 1), (777, 64, 32, 2)]   # T, D, hidden, seed


def main():
    out = {"torch_version": torch.__version__}
    for i, (T, D, H, seed) in enumerate(CASES):
        torch.manual_seed(seed)
        ft = WaveNetBlock(in_channels=D, stages=STAGES, out_dims=H, pooling=True, pooling_layers=POOL,
                          dropout_rate=0.25).eval()
        g = torch.Generator().manual_seed(400 + seed)
        feats = torch.randn(1, T, D, generator=g).abs() * 0.5
        with torch.no_grad():
            z = ft(feats.permute(0, 2, 1))
        out[f"c{i}_z"] = z[0].permute(1, 0).contiguous().numpy()
        out[f"c{i}_wsum"] = np.float64(sum(p.double().abs().sum().item() for p in ft.parameters()))
        out[f"c{i}_xsum"] = np.float64(feats.double().sum().item())
    out["cases"] = np.array([",".join(map(str, c)) for c in CASES])
    np.savez_compressed(os.path.join(HERE, "backbone_widths.npz"), **out)
    print("wrote backbone_widths.npz")


if __name__ == "__main__":
    main()
