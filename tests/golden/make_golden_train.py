"""Mints tests/golden/train.npz: loss and EVERY parameter gradient of one training step of the UNMODIFIED
reference modules -- WaveNetBlock (/root/reference/src/core/modules/temporal.py:77-147, dropout p = 0) ->
GroupNorm / ReLU / nearest interpolate / 1x1 classifier (the torch calls of src/mucon/models.py:746-773,567-582)
-> MuCon.mucon_loss (models.py:414-488, flint / box), one video at a time like trainers.py:125-131, the batch loss
being the mean over the videos -- followed by loss.backward().

Weights and features are regenerated from seeds (checksums stored); fandak / yacs are stubbed as in
make_golden_loss.py (only class names are needed)."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_loss as mgl  # noqa: E402  (installs the stubs, imports the reference MuCon)

from core.modules.temporal import WaveNetBlock  # noqa: E402

STAGES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
POOL = [1, 2, 4, 8]
D, H, C, SEED = 256, 128, 48, 7
TS = [700, 333, 1200, 130]
NS = [6, 4, 9, 3]


def build():
    torch.manual_seed(SEED)
    ft = WaveNetBlock(in_channels=D, stages=STAGES, out_dims=H, pooling=True, pooling_type="max", pooling_layers=POOL,
                      leaky=False, dropout_rate=0.0)
    gn = nn.GroupNorm(num_groups=32, num_channels=H)
    cls = nn.Conv1d(H, C, kernel_size=1)
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5)
        gn.bias.uniform_(-0.5, 0.5)
    return ft, gn, cls


def inputs():
    g = torch.Generator().manual_seed(1000 + SEED)
    feats = [torch.randn(1, T, D, generator=g).abs() * 0.5 for T in TS]
    rng = np.random.default_rng(SEED)
    lengths = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)) for n in NS]
    trs = [torch.from_numpy(rng.integers(0, C, n)).long() for n in NS]
    return feats, lengths, trs


def main():
    ft, gn, cls = build()
    ft.train(), gn.train(), cls.train()
    feats, lengths, trs = inputs()
    lengths = [l.requires_grad_(True) for l in lengths]
    total = 0.0
    per_video = []
    for f, l, tr in zip(feats, lengths, trs):
        T = f.shape[1]
        z = ft(f.permute(0, 2, 1))                                   # models.py:753-756
        z = F.relu(gn(z))                                            # models.py:759-764
        seg = cls(F.interpolate(z, T)).squeeze(0).permute(1, 0)      # models.py:574-580, 346-348
        loss = mgl.reference_loss(l, seg, tr, "flint", "box", 0.0, C)
        per_video.append(loss.item())
        total = total + loss / len(TS)
    total.backward()
    out = {"torch_version": torch.__version__, "loss": np.float32(total.item()), "per_video": np.array(per_video, np.float32),
           "TS": np.array(TS), "NS": np.array(NS), "dims": np.array([D, H, C, SEED])}
    for name, mod in (("ft", ft), ("ft_last_gn", gn), ("conv_classifier", cls)):
        for k, p in mod.named_parameters():
            out[f"g.{name}.{k}"] = p.grad.numpy()
    for i, l in enumerate(lengths):
        out[f"lengths{i}"], out[f"tr{i}"], out[f"glen{i}"] = l.detach().numpy(), trs[i].numpy(), l.grad.numpy()
    params = list(ft.parameters()) + list(gn.parameters()) + list(cls.parameters())
    out["wsum"] = np.float64(sum(p.double().abs().sum().item() for p in params))
    out["xsum"] = np.float64(sum(f.double().sum().item() for f in feats))
    np.savez_compressed(os.path.join(HERE, "train.npz"), **out)
    print("wrote train.npz: loss", total.item(), per_video)


if __name__ == "__main__":
    main()
