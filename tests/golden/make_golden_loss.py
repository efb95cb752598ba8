"""Mints tests/golden/loss.npz from the UNMODIFIED reference MuCon.mucon_loss /
calculate_mucon_loss_using_masks (/root/reference/src/mucon/models.py:414-525).  mucon.models imports
the un-vendored fandak / yacs packages, which are stubbed with empty placeholder modules here (only
class names are needed); scipy.signal.gaussian is aliased to its new location."""
import os
import sys
import types
import warnings

import numpy as np
import scipy.signal
import scipy.signal.windows
import torch
import torch.nn as nn

warnings.filterwarnings("ignore")
scipy.signal.gaussian = scipy.signal.windows.gaussian
HERE = os.path.dirname(os.path.abspath(__file__))


class _Any(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (object,), {})


class _Model(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        self.cfg = cfg


for name in ("fandak", "fandak.utils", "fandak.utils.torch", "fandak.core", "fandak.core.datasets", "yacs",
             "yacs.config", "fandak.utils.misc"):
    sys.modules[name] = _Any(name)
sys.modules["fandak"].Model = _Model
sys.modules["fandak.utils.torch"].tensor_to_numpy = lambda t: t.detach().cpu().numpy()
sys.path.insert(0, "/root/reference/src")
from mucon.models import MuCon  # noqa: E402

NS = types.SimpleNamespace


def reference_loss(lengths, seg, transcript, mtype, template, overlap, C):
    cfg = NS(model=NS(loss=NS(mucon=NS(type=mtype, template=template, overlap=overlap), mucon_weight_background=False)))
    fake = NS(cfg=cfg, teacher_forcing=True, num_classes=C)
    fake.calculate_mucon_loss_using_masks = types.MethodType(MuCon.calculate_mucon_loss_using_masks, fake)
    batch = NS(transcript=transcript)
    fo = NS(lengths=lengths, segmentation=seg, transcript=None)
    return MuCon.mucon_loss(fake, batch, fo)


def main():
    rng = np.random.default_rng(11)
    out = {"torch_version": torch.__version__}
    cases = [(1000, 6, 24, "flint", "box", 0.0), (800, 9, 24, "flint", "gaussian", 0.0), (777, 4, 20, "flint", "box", 0.1),
             (640, 5, 24, "arithmetic", "box", 0.0), (900, 7, 20, "arithmetic", "trapezoid", 0.05)]
    for i, (T, N, C, mtype, tmpl, ov) in enumerate(cases):
        lengths = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).requires_grad_(True)
        seg = torch.from_numpy(rng.standard_normal((T, C)).astype(np.float32) * 2).requires_grad_(True)
        tr = torch.from_numpy(rng.integers(0, C, N)).long()
        loss = reference_loss(lengths, seg, tr, mtype, tmpl, ov, C)
        loss.backward()
        out[f"c{i}_lengths"], out[f"c{i}_seg"], out[f"c{i}_tr"] = lengths.detach().numpy(), seg.detach().numpy(), tr.numpy()
        out[f"c{i}_loss"] = np.float32(loss.item())
        out[f"c{i}_glen"] = lengths.grad.numpy()
        out[f"c{i}_gseg_sum"] = np.float64(seg.grad.double().abs().sum().item())
        out[f"c{i}_gseg_rows"] = seg.grad.numpy()[::37].copy()
    out["cases"] = np.array([",".join(map(str, c)) for c in cases])
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)
    print("wrote loss.npz")


if __name__ == "__main__":
    main()
