"""Length -> soft mask generation for the mutual-consistency loss, on the GPU.

Drop-in surface (reference src/mucon/masks.py:8-74, called from MuCon.mucon_loss,
src/mucon/models.py:430-441):

    absolute = project_lengths_softmax(T, L)                  # [M]
    masks = create_masks(T, absolute, overlap=0.0, template="box")   # [M, T]

Like the reference, create_masks scales its length argument in place by (1 + 2*overlap)
(masks.py:61) and is differentiable w.r.t. it.  `align_corners` selects the sampling convention
the reference inherits from its torch version; None means "what the
installed torch does when the flag is omitted", i.e. False for torch >= 1.3.
create_masks_batch builds the masks of many videos in one launch.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

TEMPLATES = {"box": 0, "gaussian": 1, "trapezoid": 2}


def project_lengths_softmax(T, L):
    """T * softmax(L)  (masks.py:8-15)."""
    return T * torch.softmax(L, dim=0)


_WS = {}


def _launch_fwd(L, n_off, Ts, out_off, row_vid, V, n_rows, max_T, overlap, tid, align, L_scaled, out):
    st = torch.cuda.current_stream(L.device)
    if n_rows >= 64:
        # batches: the rows' geometry is computed by a first launch (a thread per row) into a workspace that is kept
        # per (device, size) -- at reference sizes (one video, a handful of rows) the single launch below is faster
        key = (str(L.device), int(n_rows))
        ws = _WS.get(key)
        if ws is None:
            if len(_WS) > 64:
                _WS.clear()
            ws = _WS[key] = torch.empty(16 * n_rows + 4, dtype=torch.float32, device=L.device)
        _lib.check(_lib.lib().mucon_masks_fwd_ws(
            _lib.ptr(L), _lib.ptr(n_off), _lib.ptr(Ts), _lib.ptr(out_off), _lib.ptr(row_vid), C.c_int(V), C.c_int(n_rows),
            C.c_int(max_T), C.c_float(overlap), C.c_int(tid), C.c_int(align), _lib.ptr(L_scaled), _lib.ptr(out),
            _lib.ptr(ws), C.c_void_p(st.cuda_stream)), "mucon_masks_fwd_ws")
        return
    _lib.check(_lib.lib().mucon_masks_fwd(
        _lib.ptr(L), _lib.ptr(n_off), _lib.ptr(Ts), _lib.ptr(out_off), _lib.ptr(row_vid), C.c_int(V), C.c_int(n_rows),
        C.c_int(max_T),
        C.c_float(overlap), C.c_int(tid), C.c_int(align), _lib.ptr(L_scaled), _lib.ptr(out),
        C.c_void_p(st.cuda_stream)), "mucon_masks_fwd")


def _launch_bwd(L, n_off, Ts, out_off, row_vid, V, n_rows, overlap, tid, align, gout, ws, gL):
    st = torch.cuda.current_stream(L.device)
    _lib.check(_lib.lib().mucon_masks_bwd(
        _lib.ptr(L), _lib.ptr(n_off), _lib.ptr(Ts), _lib.ptr(out_off), _lib.ptr(row_vid), C.c_int(V), C.c_int(n_rows),
        C.c_float(overlap), C.c_int(tid), C.c_int(align), _lib.ptr(gout), _lib.ptr(ws), _lib.ptr(gL),
        C.c_void_p(st.cuda_stream)), "mucon_masks_bwd")


class _MasksFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, L, meta, overlap, tid, align):
        n_off, Ts, out_off, row_vid, V, n_rows, max_T, total = meta
        Lc = L.detach().float().clone().contiguous()  # private copy: the caller's L is scaled in place later
        out = torch.empty(total, dtype=torch.float32, device=L.device)
        _launch_fwd(Lc, n_off, Ts, out_off, row_vid, V, n_rows, max_T, overlap, tid, align, None, out)
        ctx.save_for_backward(Lc)
        ctx.meta, ctx.args = meta, (overlap, tid, align)
        return out

    @staticmethod
    def backward(ctx, gout):
        (Lc,) = ctx.saved_tensors
        n_off, Ts, out_off, row_vid, V, n_rows, max_T, total = ctx.meta
        overlap, tid, align = ctx.args
        gout = gout.contiguous().float()
        ws = torch.empty(2 * n_rows, dtype=torch.float32, device=Lc.device)
        gL = torch.empty(n_rows, dtype=torch.float32, device=Lc.device)
        _launch_bwd(Lc, n_off, Ts, out_off, row_vid, V, n_rows, overlap, tid, align, gout, ws, gL)
        return gL, None, None, None, None


_META_CACHE = {}


def _meta(Ms, Ts, device):
    """Offset tables of a batch on the device.  Single-video shapes (the reference's call pattern: one create_masks
    per training step) are cached by (M, T, device), so a step does not pay an H2D copy and its implicit sync."""
    Ms = np.asarray(Ms, dtype=np.int64)
    Ts = np.asarray(Ts, dtype=np.int64)
    key = None
    if Ms.shape[0] == 1:
        key = (int(Ms[0]), int(Ts[0]), str(device))
        hit = _META_CACHE.get(key)
        if hit is not None:
            return hit
    out = _meta_build(Ms, Ts, device)
    if key is not None:
        if len(_META_CACHE) > 4096:
            _META_CACHE.clear()
        _META_CACHE[key] = out
    return out


def _meta_build(Ms, Ts, device):
    n_off = np.concatenate([[0], np.cumsum(Ms)]).astype(np.int32)
    sizes = Ms * Ts
    out_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    row_vid = np.repeat(np.arange(Ms.shape[0], dtype=np.int32), Ms)
    blob = np.concatenate([n_off.view(np.uint8), np.zeros((-n_off.nbytes) % 8, np.uint8),
                           Ts.astype(np.int32).view(np.uint8), np.zeros((-Ts.shape[0] * 4) % 8, np.uint8),
                           out_off[:-1].view(np.uint8), row_vid.view(np.uint8)])
    dev = torch.from_numpy(blob).to(device)
    o1 = n_off.nbytes + (-n_off.nbytes) % 8
    o2 = o1 + Ts.shape[0] * 4 + (-Ts.shape[0] * 4) % 8
    o3 = o2 + 8 * Ms.shape[0]
    d_noff = dev[:n_off.nbytes].view(torch.int32)
    d_T = dev[o1:o1 + Ts.shape[0] * 4].view(torch.int32)
    d_off = dev[o2:o3].view(torch.int64)
    d_rv = dev[o3:o3 + 4 * row_vid.shape[0]].view(torch.int32) if row_vid.shape[0] else None
    return (d_noff, d_T, d_off, d_rv, int(Ms.shape[0]), int(n_off[-1]), int(Ts.max(initial=0)), int(out_off[-1])), out_off


def _check(L, template):
    if template not in TEMPLATES:
        raise NameError(f"Invalid template name ({template})")  # masks.py:56
    if not L.is_cuda:
        raise _lib.MuconError("create_masks needs a CUDA tensor (there is no CPU fallback)")


def create_masks(T, L, overlap=0.0, template="box", align_corners=None):
    """[M, T] masks of one video.  Scales L in place by (1 + 2*overlap) like the reference."""
    _check(L, template)
    meta, _ = _meta([L.shape[0]], [T], L.device)
    out = _MasksFn.apply(L, meta, float(overlap), TEMPLATES[template], int(bool(align_corners)))
    if overlap != 0.0:
        L *= 1.0 + 2 * overlap  # the reference's side effect on its argument (masks.py:61)
    return out.view(L.shape[0], T)


def batch_meta(Ts, Ms, device):
    """Offset tables of a batch (one small H2D copy); reusable across calls with the same shapes."""
    return _meta(Ms, Ts, device)


def create_masks_batch(Ts, L, Ms, overlap=0.0, template="box", align_corners=None, meta=None):
    """Masks of V videos in one launch.  L: concatenated lengths [sum Ms] (not modified).
    Returns the flat buffer and the per-video offsets; video v is
    out[off[v]:off[v+1]].view(Ms[v], Ts[v]).  meta: optional result of batch_meta(Ts, Ms, device)."""
    _check(L, template)
    meta, out_off = meta if meta is not None else _meta(Ms, Ts, L.device)
    out = _MasksFn.apply(L, meta, float(overlap), TEMPLATES[template], int(bool(align_corners)))
    return out, out_off
