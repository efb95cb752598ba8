"""Mints tests/golden/masks.npz by running the UNMODIFIED reference mucon.masks.create_masks
(/root/reference/src/mucon/masks.py) in the build container.  The only shim is the alias of
scipy.signal.gaussian, which modern SciPy moved to scipy.signal.windows.
The reference passes no align_corners flag, so these vectors are the align_corners=False regime
of the installed torch (recorded in the file)."""
import os
import sys
import warnings

import numpy as np
import scipy.signal
import scipy.signal.windows
import torch

scipy.signal.gaussian = scipy.signal.windows.gaussian
sys.path.insert(0, "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")
from mucon.masks import create_masks, project_lengths_softmax  # noqa: E402

rng = np.random.default_rng(7)
out = {"torch_version": torch.__version__}
cases = []
for i, (T, M, ov, tmpl) in enumerate([(2000, 6, 0.0, "box"), (2000, 6, 0.1, "box"), (125, 6, 0.0, "box"),
                                      (1500, 9, 0.0, "gaussian"), (777, 4, 0.25, "trapezoid"),
                                      (3000, 12, 0.0, "box"), (64, 3, 0.0, "trapezoid")]):
    logits = torch.from_numpy(rng.standard_normal(M).astype(np.float32))
    L = project_lengths_softmax(T, logits).clone().requires_grad_(True)
    L_in = L.detach().clone()
    Lw = L * 1.0  # the reference scales its argument in place; keep the leaf intact
    masks = create_masks(T, Lw, overlap=ov, template=tmpl)
    g = torch.from_numpy(rng.standard_normal((M, T)).astype(np.float32))
    (masks * g).sum().backward()
    out[f"c{i}_L"] = L_in.numpy()
    out[f"c{i}_masks"] = masks.detach().numpy()
    out[f"c{i}_L_after"] = Lw.detach().numpy()
    out[f"c{i}_gout"] = g.numpy()
    out[f"c{i}_gradL"] = L.grad.numpy()
    cases.append((T, M, ov, tmpl))
out["cases"] = np.array([f"{T},{M},{ov},{t}" for T, M, ov, t in cases])
np.savez_compressed(os.path.join(HERE, "masks.npz"), **out)
print("wrote masks.npz with", len(cases), "cases")
