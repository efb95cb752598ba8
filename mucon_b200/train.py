"""Training step of the backbone on the GPU (config c5): forward with saved activations, backward on tcgen05.

What autograd derives for reference src/core/modules/temporal.py:43-53,128-147 (`WaveNetLayer` / `WaveNetBlock`)
when src/mucon/trainers.py:125-131 runs forward -> loss -> backward, for a packed batch of variable-length
videos (time-major rows, videos concatenated; `BackbonePlan`):

  forward   first_conv + ReLU                      mucon_gemm_tf32_bias_act
            per layer  h = relu(dilated(x) + bd)    mucon_conv_gemm_tf32_ex  (h is kept for the backward)
                       y = drop(1x1(h) + b1) + x    mucon_conv_gemm_tf32_ex  (mul = dropout mask/(1-p), residual)
                       [max_pool1d(2)]              mucon_maxpool2
            relu -> last_conv                       mucon_conv_gemm_tf32
  backward  data gradients   = the same conv GEMM with transposed tap slices and negated shifts; the ReLU
                               derivative is the epilogue's `gate`, the skip connection its `residual`
            weight gradients = mucon_wgrad_tf32: dW[tap] = dY^T . X[shifted] with TIME as the GEMM's K dimension,
                               operands read MN-major straight from the time-major activations; bias gradients ride on it
            max-pool         = mucon_maxpool2_bwd
  No gradient flows into the input features (they are pre-extracted I3D features, models.py:746-773).

The GroupNorm / ReLU / classifier / nearest-upsample tail of the model (models.py:759-768, 567-582) runs at the
pooled resolution on [sum Tz, 128] rows (1/16 of the frames) as packed torch ops under autograd
(`tail_logits_packed`); the mutual-consistency loss on top of it is `loss.mucon_loss_batch` (fused flint kernels).
Arithmetic: TF32 operands, fp32 accumulation and fp32 activations / gradients (>= the bf16 the config names).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .temporal import (BackbonePlan, WaveNetBlock, _kco, _stream, _tco, conv_gemm_rows, gemm_tf32_bias_act,
                       maxpool2_rows)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def conv_gemm_ex_rows(x, W_kco, plan, level, shifts, bias=None, residual=None, mul=None, gate=None, relu_mid=False,
                      relu_final=False):
    """mucon_conv_gemm_tf32_ex: gate(relu_final(relu_mid(sum_i x[t+shifts[i]] . W[i]^T + bias) * mul + residual))."""
    out = torch.empty_like(x)
    sh = _i32(shifts)
    _lib.check(_lib.lib().mucon_conv_gemm_tf32_ex(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(W_kco), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(mul), _lib.ptr(gate),
        _lib.ptr(plan.tiles[level]), C.c_int(plan.n_tiles[level]), C.c_int64(x.shape[0]), sh.ctypes.data_as(C.c_void_p),
        C.c_int(int(sh.size)), C.c_int(int(relu_mid)), C.c_int(int(relu_final)), _stream(x.device)),
        "mucon_conv_gemm_tf32_ex")
    return out


def wgrad_rows(dY, X, plan, level, shifts, xcols, out_offs, ldo, dW, dbias=None):
    """mucon_wgrad_tf32: dW[out_offs[j] + co*ldo + ci] += sum_t dY[t,co] * X[t+shifts[j], xcols[j]+ci]; dbias += colsum(dY)."""
    sh, xc = _i32(shifts), _i32(xcols)
    oo = np.ascontiguousarray(out_offs, dtype=np.int64)
    _lib.check(_lib.lib().mucon_wgrad_tf32(
        _lib.ptr(dY), _lib.ptr(X), C.c_int(X.shape[1]), _lib.ptr(plan.tiles[level]), C.c_int(plan.n_tiles[level]),
        C.c_int64(dY.shape[0]), sh.ctypes.data_as(C.c_void_p), xc.ctypes.data_as(C.c_void_p),
        oo.ctypes.data_as(C.c_void_p), C.c_int(int(sh.size)), C.c_int(int(ldo)), _lib.ptr(dW), _lib.ptr(dbias),
        _stream(dY.device)), "mucon_wgrad_tf32")


def maxpool2_bwd_rows(x, dy, plan, level):
    """gradient of maxpool2_rows(x, plan, level): x [rows(level), C], dy [rows(level+1), C] -> dx like x."""
    dx = torch.empty_like(x)
    _lib.check(_lib.lib().mucon_maxpool2_bwd(
        _lib.ptr(x), _lib.ptr(dy), _lib.ptr(plan.off[level]), _lib.ptr(plan.off[level + 1]), C.c_int(plan.V),
        C.c_int(plan.max_T[level + 1]), C.c_int(x.shape[1]), _lib.ptr(dx), _stream(x.device)), "mucon_maxpool2_bwd")
    return dx


def _param_list(block):
    """Parameters of a WaveNetBlock in the fixed order the autograd function uses."""
    ps = [block.first_conv.weight, block.first_conv.bias]
    for l in block.layers:
        ps += [l.dilated_conv.weight, l.dilated_conv.bias, l.conv_1x1.weight, l.conv_1x1.bias]
    ps += [block.last_conv.weight, block.last_conv.bias]
    return ps


class _WaveNetBlockFn(torch.autograd.Function):
    """forward_packed of a WaveNetBlock with the activations the backward needs kept in HBM."""

    @staticmethod
    def forward(ctx, block, plan, feats, drop_masks, *params):
        H = block.out_dims
        if H != 128 or block.in_channels % 128 != 0:
            raise NotImplementedError("the training kernels are built for 128 hidden channels and in_channels % 128 == 0")
        V = plan.V
        feats = feats.detach().contiguous().float()
        wf = block.first_conv.weight.detach()[:, :, 0].contiguous().float()
        x = gemm_tf32_bias_act(feats, wf, block.first_conv.bias.detach().float().contiguous(), relu=True)
        saved = dict(x0=x, layers=[])
        level = 0
        last = block.num_stages - 1
        if block.num_stages:
            # tensor-core layouts of every layer's weights in four launches: [k][Cout][Cin] for the forward GEMMs,
            # [k][Cin][Cout] (tap slices transposed) for the data gradients
            Wd = torch.stack([l.dilated_conv.weight.detach().float() for l in block.layers])          # [L, co, ci, 3]
            W1 = torch.stack([l.conv_1x1.weight.detach().float()[:, :, 0] for l in block.layers])     # [L, co, ci]
            kco_d, kco_1 = Wd.permute(0, 3, 1, 2).contiguous(), W1
            saved["tco_d"], saved["tco_1"] = Wd.permute(0, 3, 2, 1).contiguous(), W1.transpose(1, 2).contiguous()
        for i, l in enumerate(block.layers):
            d = block.stages[i]
            pooled = block.pooling and i in block.pooling_layers
            h = conv_gemm_ex_rows(x, kco_d[i].view(3 * H, H), plan, level, (-d, 0, d),
                                  bias=l.dilated_conv.bias.detach().float().contiguous(), relu_mid=True)
            y = conv_gemm_ex_rows(h, kco_1[i], plan, level, (0,),
                                  bias=l.conv_1x1.bias.detach().float().contiguous(), mul=drop_masks[i], residual=x,
                                  relu_final=(i == last and not pooled))
            rec = dict(x=x, h=h, level=level, pooled=pooled, y=None)
            if pooled:
                rec["y"] = y
                x = maxpool2_rows(y, plan, level)
                level += 1
                if i == last:
                    x = torch.relu_(x)
            else:
                x = y
            saved["layers"].append(rec)
        if block.num_stages == 0:
            x = torch.relu(x)
        saved["x_last"] = x       # relu(.) of the last layer's output: input of last_conv and its own ReLU gate
        saved["level_last"] = level
        out = conv_gemm_rows(x, _kco(block.last_conv), block.last_conv.bias.detach().float().contiguous(), plan, level)
        ctx.block, ctx.plan, ctx.saved, ctx.feats, ctx.drop_masks = block, plan, saved, feats, drop_masks
        return out

    @staticmethod
    def backward(ctx, dout):
        block, plan, saved, feats, drop_masks = ctx.block, ctx.plan, ctx.saved, ctx.feats, ctx.drop_masks
        dev = dout.device
        H, D = block.out_dims, block.in_channels
        L = block.num_stages
        # one zeroed buffer for every gradient (the wgrad kernel accumulates with red.global.add)
        sizes = [H * D, H] + [3 * H * H, H, H * H, H] * L + [H * H, H]
        flat = torch.zeros(int(sum(sizes)), dtype=torch.float32, device=dev)
        views, o = [], 0
        for s in sizes:
            views.append(flat[o:o + s])
            o += s
        g_first_w, g_first_b = views[0], views[1]
        g_last_w, g_last_b = views[-2], views[-1]
        dout = dout.contiguous().float()
        lvl = saved["level_last"]
        # last_conv (temporal.py:145) and the ReLU in front of it (:144)
        wgrad_rows(dout, saved["x_last"], plan, lvl, (0,), (0,), (0,), H, g_last_w, g_last_b)
        dx = conv_gemm_ex_rows(dout, _tco(block.last_conv).view(H, H), plan, lvl, (0,), gate=saved["x_last"])
        for i in reversed(range(L)):
            l, rec = block.layers[i], saved["layers"][i]
            d, level = block.stages[i], rec["level"]
            g_wd, g_bd, g_w1, g_b1 = views[2 + 4 * i: 6 + 4 * i]
            dy = maxpool2_bwd_rows(rec["y"], dx, plan, level) if rec["pooled"] else dx          # temporal.py:137-139
            dym = dy * drop_masks[i] if drop_masks[i] is not None else dy                        # temporal.py:51
            wgrad_rows(dym, rec["h"], plan, level, (0,), (0,), (0,), H, g_w1, g_b1)             # conv_1x1
            dh = conv_gemm_ex_rows(dym, saved["tco_1"][i], plan, level, (0,), gate=rec["h"])    # ReLU of :49
            wgrad_rows(dh, rec["x"], plan, level, (-d, 0, d), (0, 0, 0), (0, H * H, 2 * H * H), H, g_wd, g_bd)
            # dx = dy (skip connection) + dilated^T(dh); the ReLU behind first_conv gates layer 0's result
            dx = conv_gemm_ex_rows(dh, saved["tco_d"][i].view(3 * H, H), plan, level, (d, 0, -d), residual=dy,
                                   gate=saved["x0"] if i == 0 else None)
        if L == 0:
            dx = dx * (saved["x0"] > 0)
        nb = D // 128
        wgrad_rows(dx, feats, plan, 0, [0] * nb, [128 * j for j in range(nb)], [128 * j for j in range(nb)], D,
                   g_first_w, g_first_b)
        grads = [g_first_w.view(H, D, 1), g_first_b]
        for i in range(L):
            g_wd, g_bd, g_w1, g_b1 = views[2 + 4 * i: 6 + 4 * i]
            grads += [g_wd.view(3, H, H).permute(1, 2, 0), g_bd, g_w1.view(H, H, 1), g_b1]
        grads += [g_last_w.view(H, H, 1), g_last_b]
        ctx.saved = None
        return (None, None, None, None) + tuple(grads)


def wavenet_forward_train_packed(block: WaveNetBlock, feats, plan: BackbonePlan, generator=None):
    """WaveNetBlock.forward (temporal.py:128-147) on packed rows with autograd through the tcgen05 kernels.
    In training mode every layer's dropout (temporal.py:51) draws an inverted-dropout mask with torch's generator."""
    if block.leaky or (block.pooling and block.pooling_type != "max"):
        raise NotImplementedError("leaky ReLU / avg pooling")
    if not feats.is_cuda:
        raise _lib.MuconError("the backbone needs CUDA tensors (there is no CPU fallback)")
    p = block.dropout_rate if block.training else 0.0
    levels, level = [], 0
    for i in range(block.num_stages):
        levels.append(level)
        if block.pooling and i in block.pooling_layers:
            level += 1
    masks = [None] * block.num_stages
    if p > 0 and block.num_stages:
        # every layer's inverted-dropout mask from one draw (two launches instead of three per layer)
        sizes = [plan.rows[lv] * block.out_dims for lv in levels]
        flat = torch.empty(int(sum(sizes)), dtype=torch.float32, device=feats.device)
        flat.bernoulli_(1.0 - p, generator=generator).mul_(1.0 / (1.0 - p))
        o = 0
        for i, n in enumerate(sizes):
            masks[i] = flat[o:o + n].view(plan.rows[levels[i]], block.out_dims)
            o += n
    return _WaveNetBlockFn.apply(block, plan, feats, masks, *_param_list(block))


# ------------------------------------------------------------------------------------------------------------------
# tail at the pooled resolution (packed torch ops under autograd)
def _plan_train_tables(plan, device):
    """row -> video table at the pooled resolution and the nearest-neighbour frame -> pooled-row index
    (F.interpolate(mode="nearest"): min(floor(t * float32(Tz / T)), Tz - 1), models.py:577)."""
    if getattr(plan, "_train_tables", None) is None:
        Tz, T = plan.T[-1], plan.T[0]
        vid = np.repeat(np.arange(plan.V), Tz)
        idx = []
        off_z = plan.off_host[-1]
        for v in range(plan.V):
            if T[v] == 0:
                continue
            scale = np.float32(Tz[v]) / np.float32(T[v])
            i = np.floor(np.arange(T[v], dtype=np.float32) * scale).astype(np.int64)
            idx.append(np.minimum(i, max(int(Tz[v]) - 1, 0)) + off_z[v])
        idx = np.concatenate(idx) if idx else np.zeros(0, np.int64)
        plan._train_tables = (torch.from_numpy(vid.astype(np.int64)).to(device), torch.from_numpy(idx).to(device),
                              torch.from_numpy(Tz.astype(np.float32)).to(device))
    return plan._train_tables


def groupnorm_packed(x, vid, counts, V, weight, bias, groups, eps):
    """nn.GroupNorm over (time x channels of the group) per video, on packed [rows, C] activations (models.py:759-764)."""
    R, Cc = x.shape
    cpg = Cc // groups
    xg = x.view(R, groups, cpg)
    n = (counts * cpg).clamp_min(1.0)[:, None]
    mean = torch.zeros((V, groups), dtype=x.dtype, device=x.device).index_add_(0, vid, xg.sum(-1)) / n
    xc = xg - mean[vid][:, :, None]
    var = torch.zeros((V, groups), dtype=x.dtype, device=x.device).index_add_(0, vid, (xc * xc).sum(-1)) / n
    xh = xc * torch.rsqrt(var + eps)[vid][:, :, None]
    return xh.reshape(R, Cc) * weight[None, :] + bias[None, :]


class _GroupNormReluFn(torch.autograd.Function):
    """relu(GroupNorm(x)) per video on packed rows: mucon_groupnorm_relu / mucon_groupnorm_relu_bwd."""

    @staticmethod
    def forward(ctx, x, gamma, beta, plan, groups, eps, relu):
        from .temporal import groupnorm_relu_rows
        lvl = len(plan.off) - 1
        xc, g, b = x.detach().contiguous().float(), gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        ctx.save_for_backward(xc, g, b)
        ctx.args = (plan, lvl, groups, eps, relu)
        return groupnorm_relu_rows(xc, g, b, plan.off[lvl], plan.V, groups, eps, relu)

    @staticmethod
    def backward(ctx, dy):
        xc, g, b = ctx.saved_tensors
        plan, lvl, groups, eps, relu = ctx.args
        dy = dy.contiguous().float()
        dx = torch.empty_like(xc)
        dgb = torch.zeros((2, xc.shape[1]), dtype=torch.float32, device=xc.device)
        _lib.check(_lib.lib().mucon_groupnorm_relu_bwd(
            _lib.ptr(xc), _lib.ptr(dy), _lib.ptr(g), _lib.ptr(b), _lib.ptr(plan.off[lvl]), C.c_int(plan.V),
            C.c_int(xc.shape[1]), C.c_int(groups), C.c_float(eps), C.c_int(int(relu)), _lib.ptr(dx), _lib.ptr(dgb[0]),
            _lib.ptr(dgb[1]), _stream(xc.device)), "mucon_groupnorm_relu_bwd")
        return dx, dgb[0], dgb[1], None, None, None, None


class _ExpandRowsFn(torch.autograd.Function):
    """nearest-neighbour expansion of pooled-resolution rows to every frame: mucon_expand_rows / _bwd."""

    @staticmethod
    def forward(ctx, table, plan):
        lvl = len(plan.off) - 1
        tb = table.detach().contiguous().float()
        out = torch.empty((plan.rows[0], tb.shape[1]), dtype=torch.float32, device=tb.device)
        _lib.check(_lib.lib().mucon_expand_rows(
            _lib.ptr(tb), _lib.ptr(plan.off[lvl]), _lib.ptr(plan.off[0]), C.c_int(plan.V), C.c_int(plan.max_T[0]),
            C.c_int(tb.shape[1]), _lib.ptr(out), _stream(tb.device)), "mucon_expand_rows")
        ctx.plan, ctx.shape = plan, tb.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        plan = ctx.plan
        lvl = len(plan.off) - 1
        gout = gout.contiguous().float()
        gt = torch.empty(ctx.shape, dtype=torch.float32, device=gout.device)
        _lib.check(_lib.lib().mucon_expand_rows_bwd(
            _lib.ptr(gout), _lib.ptr(plan.off[lvl]), _lib.ptr(plan.off[0]), C.c_int(plan.V), C.c_int(plan.max_T[lvl]),
            C.c_int(gout.shape[1]), _lib.ptr(gt), _stream(gout.device)), "mucon_expand_rows_bwd")
        return gt, None


def tail_logits_packed(model, z_pre, plan, fused=True):
    """ft output [sum Tz, H] -> frame logits [sum T, classes]: GroupNorm + ReLU (models.py:759-768), 1x1 classifier at
    the pooled resolution and nearest-neighbour expansion (models.py:567-582; the two commute).  fused=True runs
    GroupNorm(+ReLU) and the expansion, forward and backward, as CUDA kernels; False as packed torch ops."""
    z = z_pre
    if fused and model.hidden_size == 128 and z_pre.is_cuda:
        if model.last_gn:
            z = _GroupNormReluFn.apply(z, model.ft_last_gn.weight, model.ft_last_gn.bias, plan,
                                       model.ft_last_gn.num_groups, model.ft_last_gn.eps, model.last_relu)
        elif model.last_relu:
            z = torch.relu(z)
        logits_z = torch.addmm(model.conv_classifier.bias, z, model.conv_classifier.weight[:, :, 0].t())
        return _ExpandRowsFn.apply(logits_z, plan), z
    vid, idx, counts = _plan_train_tables(plan, z_pre.device)
    if model.last_gn:
        z = groupnorm_packed(z, vid, counts, plan.V, model.ft_last_gn.weight, model.ft_last_gn.bias,
                             model.ft_last_gn.num_groups, model.ft_last_gn.eps)
    if model.last_relu:
        z = torch.relu(z)
    logits_z = z @ model.conv_classifier.weight[:, :, 0].t() + model.conv_classifier.bias[None, :]
    return logits_z[idx], z


def forward_train_packed(model, feats, plan, generator=None):
    """MuCon.temporal_modeling_forward + frame_classifier_forward for a packed batch, differentiable w.r.t. every
    backbone / GroupNorm / classifier parameter.  Returns (frame logits [sum T, classes], z [sum Tz, H])."""
    z_pre = wavenet_forward_train_packed(model.ft, feats, plan, generator=generator)
    return tail_logits_packed(model, z_pre, plan)


class TrainStep:
    """One training step of the backbone path for a fixed batch shape, captured in a CUDA graph: forward (tcgen05
    kernels above) -> GroupNorm / classifier tail -> batched flint loss -> backward -> optimizer step, replayed with
    one launch per step (trainers.py:125-131 does the same four calls per video from Python).

    The step is launch-bound at reference batch sizes (about 200 small kernels for 32 videos), which is what the
    graph removes.  Inputs live in static buffers: copy the next batch into `feats`, `lengths`, `transcripts` (same
    Ts / Ms) and call `run()`; the loss is read from `loss` (a device scalar)."""

    def __init__(self, model, Ts, Ms, optimizer=None, template="box", overlap=0.0, graph=True, warmup=3):
        from .loss import _flint_meta, mucon_loss_batch
        self.model, self.optimizer = model, optimizer
        dev = model.conv_classifier.weight.device
        self.Ts, self.Ms = [int(t) for t in Ts], [int(m) for m in Ms]
        self.plan = model.plan(self.Ts, dev)
        self.meta = _flint_meta(self.Ms, self.Ts, dev)
        D = model.ft.in_channels
        self.feats = torch.zeros((int(sum(self.Ts)), D), dtype=torch.float32, device=dev)
        self.lengths = torch.zeros(int(sum(self.Ms)), dtype=torch.float32, device=dev, requires_grad=True)
        self.transcripts = torch.zeros(int(sum(self.Ms)), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self._loss_fn = lambda seg: mucon_loss_batch(self.lengths, seg, self.transcripts, self.Ms, self.Ts,
                                                     template=template, overlap=overlap, meta=self.meta)
        self.graph = None
        self._use_graph, self._warmup = graph, warmup

    def _eager(self):
        for p in self.model.parameters():
            p.grad = None
        self.lengths.grad = None
        seg, _ = forward_train_packed(self.model, self.feats, self.plan)
        loss = self._loss_fn(seg)
        loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        self.loss.copy_(loss.detach())

    def run(self):
        if not self._use_graph:
            self._eager()
            return self.loss
        if self.graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up off the capture stream (allocator, lazy tables, autograd)
                for _ in range(self._warmup):
                    self._eager()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._eager()
        self.graph.replay()
        return self.loss
