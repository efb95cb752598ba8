"""Mask generation for the mutual-consistency loss -- ORACLE / TEST INFRASTRUCTURE ONLY.

Restates reference src/mucon/masks.py:8-120.

create_masks_torch      the same library calls as the reference (cumsum -> affine parameters ->
                        F.affine_grid -> F.grid_sample of a 100-tap template), written from the
                        formulas, differentiable w.r.t. L, `align_corners` exposed (the reference
                        passes neither flag: torch 1.1 in its docker meant True, torch >= 1.3 means
                        False -- SURVEY.md section 0.7).
create_masks_closed_form  the closed form of that composition in NumPy float64:
                          align_corners=False: u = (t + 0.5 - pi)*W/Ls - 0.5
                          align_corners=True : u = (t*T/(T-1) - pi)*(W-1)/Ls
                          out = lerp(template, u), taps outside [0, W) are zero.
templates               box / gaussian(std = W/5) / trapezoid (masks.py:34-54).
"""
import numpy as np
import torch
import torch.nn.functional as F

W = 100  # TEMPLATE_WIDTH, masks.py:32


def template(name):
    if name == "box":
        return np.ones(W, dtype=np.float32)
    if name == "gaussian":  # scipy.signal.gaussian(M=100, std=20) == exp(-0.5*((n-(M-1)/2)/std)^2)
        n = np.arange(W, dtype=np.float64) - (W - 1) / 2.0
        return np.exp(-0.5 * (n / (W / 5)) ** 2).astype(np.float32)
    if name == "trapezoid":
        t = torch.ones(W)
        w1 = W / 2
        t[: int(w1 / 2)] = torch.arange(start=0.5, end=1, step=(1 - 0.5) / (w1 / 2))
        t[-int(w1 / 2):] = torch.arange(start=1, end=0.5, step=(0.5 - 1) / (w1 / 2))
        return t.numpy().astype(np.float32)
    raise NameError(f"Invalid template name ({name})")


def project_lengths_softmax(T, L):
    return T * torch.softmax(L, dim=0)  # masks.py:8-15


def create_masks_torch(T, L, overlap=0.0, template_name="box", align_corners=False):
    """L: [M] float32 tensor (not modified).  Returns ([M, T] masks, scaled lengths)."""
    M = L.shape[0]
    tmpl = torch.from_numpy(template(template_name)).to(L.device).repeat(M, 1).view(M, 1, 1, W)
    pis = torch.cumsum(L, 0) - L
    Ls = L * (1.0 + 2 * overlap)
    pis = pis - Ls * (overlap / 2)
    s = T / Ls
    x = (pis + Ls / 2 - T / 2) / (-(Ls / 2))
    zero = torch.zeros_like(s)
    theta = torch.stack([torch.stack([s, zero, x], 1), torch.stack([zero, s, zero], 1)], 1)  # [[s,0,x],[0,s,0]]
    grid = F.affine_grid(theta, torch.Size((M, 1, 1, T)), align_corners=align_corners)
    out = F.grid_sample(tmpl, grid, mode="bilinear", padding_mode="zeros", align_corners=align_corners)
    return out.view(M, T), Ls


def create_masks_closed_form(T, L, overlap=0.0, template_name="box", align_corners=False):
    L = np.asarray(L, dtype=np.float64)
    tm = template(template_name).astype(np.float64)
    pis = np.cumsum(L) - L
    Ls = L * (1.0 + 2 * overlap)
    pis = pis - Ls * (overlap / 2)
    t = np.arange(T, dtype=np.float64)[None, :]
    if align_corners:
        u = (t * (T / (T - 1.0) if T > 1 else 0.0) - pis[:, None]) * (W - 1) / Ls[:, None]
    else:
        u = (t + 0.5 - pis[:, None]) * W / Ls[:, None] - 0.5
    fl = np.floor(u)
    w1 = u - fl
    i0 = fl.astype(np.int64)
    pad = np.concatenate([[0.0], tm, [0.0]])  # index -1 -> 0, index W -> W+1
    a = pad[np.clip(i0 + 1, 0, W + 1)]
    b = pad[np.clip(i0 + 2, 0, W + 1)]
    return (a * (1 - w1) + b * w1).astype(np.float32)
