"""Mints tests/golden/backbone_alt.npz: outputs of the UNMODIFIED reference MSTCNPPFirstStage and
NoFt modules (/root/reference/src/core/modules/temporal.py:150-204, :56-74) in eval mode.  Weights and
inputs are regenerated from seeds; the file keeps checksums of both (see make_golden_backbone.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))
from core.modules.temporal import MSTCNPPFirstStage, NoFt  # noqa: E402

CASES = [("mstcnpp", 11, 256, 700, 0), ("mstcnpp", 11, 2048, 333, 1), ("mstcnpp", 6, 64, 1500, 2),
         ("noft", 0, 256, 500, 3), ("noft", 0, 2048, 129, 4)]


def build(kind, L, D, seed):
    torch.manual_seed(seed)
    if kind == "mstcnpp":
        return MSTCNPPFirstStage(num_layers=L, num_f_maps=128, input_dim=D, output_dim=128).eval()
    return NoFt(in_chnnels=D, out_dims=128).eval()


def main():
    out = {"torch_version": torch.__version__}
    for i, (kind, L, D, T, seed) in enumerate(CASES):
        m = build(kind, L, D, seed)
        g = torch.Generator().manual_seed(200 + seed)
        feats = torch.randn(1, T, D, generator=g).abs() * 0.5
        with torch.no_grad():
            z = m(feats.permute(0, 2, 1))
        out[f"c{i}_z"] = z[0].permute(1, 0).contiguous().numpy()   # [T', 128]
        out[f"c{i}_wsum"] = np.float64(sum(p.double().abs().sum().item() for p in m.parameters()))
        out[f"c{i}_xsum"] = np.float64(feats.double().sum().item())
    out["cases"] = np.array([",".join(map(str, c)) for c in CASES])
    np.savez_compressed(os.path.join(HERE, "backbone_alt.npz"), **out)
    print("wrote backbone_alt.npz")


if __name__ == "__main__":
    main()
