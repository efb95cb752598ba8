"""CPU: mask oracle restatements against the reference's frozen outputs (tests/golden/masks.npz,
minted by tests/golden/make_golden_masks.py from the unmodified reference)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import coracle
from oracle import masks as omasks
from tests.util import mask_atol

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "masks.npz"))
CASES = [c.split(",") for c in G["cases"]]


def case(i):
    T, M, ov, tmpl = CASES[i]
    return int(T), int(M), float(ov), tmpl, {k: G[f"c{i}_{k}"] for k in ("L", "masks", "L_after", "gout", "gradL")}


@pytest.mark.parametrize("i", range(len(CASES)))
def test_torch_restatement_matches_reference(i):
    T, M, ov, tmpl, g = case(i)
    L = torch.from_numpy(g["L"].copy()).requires_grad_(True)
    masks, Ls = omasks.create_masks_torch(T, L, ov, tmpl, align_corners=False)
    assert np.abs(masks.detach().numpy() - g["masks"]).max() <= 1e-6
    assert np.allclose(Ls.detach().numpy(), g["L_after"], rtol=1e-6)
    (masks * torch.from_numpy(g["gout"])).sum().backward()
    assert np.allclose(L.grad.numpy(), g["gradL"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_closed_form_and_c_port_match_reference(i):
    """The reference evaluates the template coordinate in float32 as u = ((g*s + x + 1)*W - 1)/2 with
    |g*s + x| up to s = T/L, so u carries an error of about (T/L) * W/2 * 2^-23 and a ramp value
    moves by as much.  Tolerance: 3e-5 * max(1, T / min L) (tests/util.mask_atol)."""
    T, M, ov, tmpl, g = case(i)
    atol = mask_atol(T, g["L"])
    cf = omasks.create_masks_closed_form(T, g["L"], ov, tmpl, align_corners=False)
    assert np.abs(cf - g["masks"]).max() <= atol
    cm, Ls = coracle.masks(g["L"], T, ov, omasks.template(tmpl), False)
    assert np.abs(cm - g["masks"]).max() <= atol
    assert np.allclose(Ls, g["L_after"], rtol=1e-6)


def test_box_masks_are_one_inside_their_segment_and_zero_outside():
    rng = np.random.default_rng(0)
    T = 1234
    L = rng.dirichlet(np.ones(7)) * T
    m = omasks.create_masks_closed_form(T, L, 0.0, "box", False)
    ends = np.cumsum(L)
    starts = ends - L
    t = np.arange(T) + 0.5
    for i in range(7):
        ramp = L[i] / 100 + 1
        assert np.all(m[i][(t > starts[i] + ramp) & (t < ends[i] - ramp)] == 1.0)
        assert np.all(m[i][(t < starts[i] - ramp) | (t > ends[i] + ramp)] == 0.0)


def test_library_templates_match_oracle():
    from mucon_b200 import _lib, build
    build.build()
    lib = _lib.lib()
    for name, tid in (("box", 0), ("gaussian", 1), ("trapezoid", 2)):
        out = np.zeros(100, dtype=np.float32)
        assert lib.mucon_mask_template_h(tid, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.abs(out - omasks.template(name)).max() <= 1e-7, name


def test_bad_template_name_raises_nameerror():
    with pytest.raises(NameError):  # masks.py:56
        omasks.template("triangle")
