"""Dilated temporal-conv backbone on the GPU -- host side (inference).

Drop-in surface: `WaveNetBlock` with the constructor, parameter names and forward signature of
reference src/core/modules/temporal.py:77-147 (`first_conv`, `l_{i}.dilated_conv`, `l_{i}.conv_1x1`,
`last_conv`), so a reference state_dict loads unchanged, the alternates `MSTCNPPFirstStage` (:150-204)
and `NoFt` (:56-74) selected by `model.ft.type`, and `MuConBackbone` with the attribute
names the reference model uses around it (`ft`, `ft_last_gn`, `conv_classifier`,
src/mucon/models.py:160-191,276-278) and its three forward helpers (models.py:360-374,567-582,746-773).

Internally activations are time-major ([rows, channels], videos concatenated) and every op is a
call into libmucon_b200.so (include/mucon_b200.h); `BackbonePlan` holds the per-resolution row
offsets of a batch of variable-length videos.  This module is the inference forward; the training step (forward with
kept activations, dropout, backward on tcgen05) is mucon_b200/train.py, the s-head mucon_b200/shead.py, the whole
test-time pipeline mucon_b200/inference.py.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class BackbonePlan:
    """Row offsets of a batch of videos at every pooling level (floor halving per max-pool)."""

    def __init__(self, T, n_pools, device):
        self.device = torch.device(device)
        T = np.asarray(T, dtype=np.int64)
        self.V = int(T.shape[0])
        self.T = [T]
        for _ in range(n_pools):
            self.T.append(self.T[-1] // 2)
        offs = [np.concatenate([[0], np.cumsum(t)]).astype(np.int64) for t in self.T]
        self.rows = [int(o[-1]) for o in offs]
        self.max_T = [int(t.max(initial=0)) for t in self.T]
        self.off_host = offs
        flat = torch.from_numpy(np.concatenate(offs)).to(self.device)
        n = self.V + 1
        self.off = [flat[i * n:(i + 1) * n] for i in range(len(offs))]
        # 128-row tiles of every video at every level, for the tcgen05 conv kernel:
        # {int64 row0, int32 t0, int32 T} per tile, longest videos first
        self.tiles, self.n_tiles, self.tile_vid = [], [], []
        for lvl, t in enumerate(self.T):
            nt = (t + 127) // 128
            vid = np.repeat(np.arange(self.V), nt)
            first = np.concatenate([[0], np.cumsum(nt)])[:-1]
            t0 = (np.arange(int(nt.sum())) - np.repeat(first, nt)) * 128
            rec = np.zeros(int(nt.sum()), dtype=[("row0", "<i8"), ("t0", "<i4"), ("T", "<i4")])
            rec["row0"], rec["t0"], rec["T"] = offs[lvl][:-1][vid], t0, t[vid]
            self.n_tiles.append(int(rec.shape[0]))
            self.tile_vid.append(torch.from_numpy(vid.astype(np.int32)).to(self.device) if rec.shape[0]
                                 else torch.zeros(1, dtype=torch.int32, device=self.device))
            self.tiles.append(torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).to(self.device)
                              if rec.shape[0] else torch.zeros(16, dtype=torch.uint8, device=self.device))
        # tiles of the fused layer kernel: {int64 row0, int64 row0_out, int32 t0, int32 T}, for a layer
        # that keeps the resolution ("same") or is followed by a max-pool ("pool")
        self.ltiles = {}
        for lvl, t in enumerate(self.T):
            nt = (t + 127) // 128
            vid = np.repeat(np.arange(self.V), nt)
            first = np.concatenate([[0], np.cumsum(nt)])[:-1]
            t0 = (np.arange(int(nt.sum())) - np.repeat(first, nt)) * 128
            # the paired kernel walks tiles two at a time, both from the same video: every video is padded
            # to an even number of tiles (a padding tile starts at t0 = 128 * tiles >= T: nothing is stored)
            nt2 = (nt + 1) // 2 * 2
            vid2 = np.repeat(np.arange(self.V), nt2)
            first2 = np.concatenate([[0], np.cumsum(nt2)])[:-1]
            t02 = (np.arange(int(nt2.sum())) - np.repeat(first2, nt2)) * 128
            for kind in ("same", "pool"):
                if kind == "pool" and lvl + 1 >= len(self.T):
                    continue
                for suffix, (vv, tt0) in (("", (vid, t0)), ("2", (vid2, t02))):
                    rec = np.zeros(int(vv.shape[0]), dtype=[("row0", "<i8"), ("row0_out", "<i8"), ("t0", "<i4"), ("T", "<i4")])
                    rec["row0"], rec["t0"], rec["T"] = offs[lvl][:-1][vv], tt0, t[vv]
                    rec["row0_out"] = offs[lvl + 1][:-1][vv] if kind == "pool" else rec["row0"]
                    self.ltiles[(lvl, kind + suffix)] = (
                        torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).to(self.device)
                        if rec.shape[0] else torch.zeros(24, dtype=torch.uint8, device=self.device), int(rec.shape[0]))


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def gemm_tf32_bias_act(A, W, bias, relu, out_bf16=False, out_dtype=None):
    """act(A @ W.T + bias) on tcgen05 (TF32 operands, fp32 accumulate).  A [M,K], W [128,K].
    out_dtype: torch.float32 (default), torch.bfloat16 or torch.float16 (the input of the 16-bit layer kernel)."""
    out_dtype = out_dtype or (torch.bfloat16 if out_bf16 else torch.float32)
    out = torch.empty((A.shape[0], W.shape[0]), dtype=out_dtype, device=A.device)
    if out_dtype == torch.float32:
        _lib.check(_lib.lib().mucon_gemm_tf32_bias_act(
            _lib.ptr(A), C.c_int64(A.shape[0]), C.c_int(A.shape[1]), _lib.ptr(W), C.c_int(W.shape[0]), _lib.ptr(bias),
            _lib.ptr(out), C.c_int(int(relu)), _stream(A.device)), "mucon_gemm_tf32_bias_act")
    else:
        _lib.check(_lib.lib().mucon_gemm_tf32_bias_act_bf16(
            _lib.ptr(A), C.c_int64(A.shape[0]), C.c_int(A.shape[1]), _lib.ptr(W), C.c_int(W.shape[0]), _lib.ptr(bias),
            _lib.ptr(out), C.c_int(int(relu)), C.c_int(int(out_dtype == torch.float16)), _stream(A.device)),
            "mucon_gemm_tf32_bias_act_bf16")
    return out


def conv1d_rows(x, W_tco, bias, off, V, max_T, dilation=1, relu_in=False, relu_out=False, residual=None):
    """k=1/k=3 dilated conv over time-major rows.  W_tco: [taps, Cin, Cout]."""
    taps, Cin, Cout = W_tco.shape
    out = torch.empty((x.shape[0], Cout), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().mucon_conv1d(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(W_tco), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(off), C.c_int(V),
        C.c_int(max_T), C.c_int(Cin), C.c_int(Cout), C.c_int(taps), C.c_int(dilation), C.c_int(int(relu_in)),
        C.c_int(int(relu_out)), _stream(x.device)), "mucon_conv1d")
    return out


def conv_gemm_rows(x, W_kco, bias, plan, level, dilation=1, relu_mid=False, relu_final=False, residual=None):
    """128 -> 128 channel conv (k = 1 / 3) on tcgen05.  W_kco: [taps*128, 128] = weight [k][Cout][Cin]."""
    out = torch.empty_like(x)
    taps = W_kco.shape[0] // 128
    _lib.check(_lib.lib().mucon_conv_gemm_tf32(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(W_kco), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(plan.tiles[level]),
        C.c_int(plan.n_tiles[level]), C.c_int64(x.shape[0]), C.c_int(taps), C.c_int(dilation), C.c_int(int(relu_mid)),
        C.c_int(int(relu_final)), _stream(x.device)), "mucon_conv_gemm_tf32")
    return out


# Cluster pairs sharing the weight traffic through TMA multicast (mucon_wavenet_layer_tf32_pair).  Bit-identical
# results, but measured slower on B200 (level-0 layer 2.04 vs 1.91 ms, whole encode 14.5 vs 13.0 ms on the
# 1712-video batch): the layer kernel is bound by its per-tile GEMM -> epilogue -> GEMM -> epilogue chain, not
# by L2 -> shared-memory weight reads, and the lockstep between the two CTAs adds to that chain.  Off by default.
LAYER_PAIRS = os.environ.get("MUCON_LAYER_PAIRS", "0") != "0"


def wavenet_layer_rows(x, Wd_kco, bd, W1_kco, b1, plan, level, dilation, pool, relu_final, pair=None):
    """One WaveNet layer (+ optional max-pool) in a single tcgen05 launch.  x [rows(level), 128]."""
    pair = LAYER_PAIRS if pair is None else pair
    tiles, n_tiles = plan.ltiles[(level, ("pool" if pool else "same") + ("2" if pair else ""))]
    out = torch.empty((plan.rows[level + 1] if pool else x.shape[0], 128), dtype=torch.float32, device=x.device)
    fn = _lib.lib().mucon_wavenet_layer_tf32_pair if pair else _lib.lib().mucon_wavenet_layer_tf32
    _lib.check(fn(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(Wd_kco), _lib.ptr(bd), _lib.ptr(W1_kco), _lib.ptr(b1), _lib.ptr(tiles),
        C.c_int(n_tiles), C.c_int64(x.shape[0]), C.c_int(dilation), C.c_int(int(pool)), C.c_int(int(relu_final)),
        _stream(x.device)), "mucon_wavenet_layer_tf32")
    return out


# Arithmetic of the 128 -> 128 layers: "fp16" / "bf16" (16-bit activations and weights, tcgen05 kind::f16, weights
# resident in shared memory: the fast path; fp16's 11-bit mantissa keeps the residual stream as accurate as TF32,
# bf16 has fp32's range but 8x the rounding error), "tf32" (fp32 activations read as TF32) or "fp32" (CUDA cores,
# exact fp32).
DEFAULT_PRECISION = os.environ.get("MUCON_BACKBONE_PRECISION", "fp16")


_ZERO_BIAS = np.zeros(128, dtype=np.float32)


def _host_f32(t):
    """contiguous float32 host copy of a (small) tensor, as a ctypes pointer keeps it alive through the call"""
    a = np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32) if isinstance(t, torch.Tensor) else \
        np.ascontiguousarray(t, dtype=np.float32)
    return a


def wavenet_layer_bf16_rows(x, Wd_kco16, bd, W1_kco16, b1, plan, level, dilation, pool, relu_final, out_f32=False,
                            residual=True):
    """One WaveNet layer (+ optional max-pool) in a single tcgen05 launch, 16-bit operands.  x [rows(level), 128]
    bfloat16 or float16 (the weights must have the same type).
    bd / b1: biases, as host float32 arrays (numpy) or tensors (copied to the host: pass numpy in hot loops).
    residual=False drops the skip connection (out = conv_1x1(relu(conv(x) + bd)) + b1)."""
    tiles, n_tiles = plan.ltiles[(level, "pool" if pool else "same")]
    assert x.dtype in (torch.bfloat16, torch.float16) and Wd_kco16.dtype == x.dtype and W1_kco16.dtype == x.dtype
    out = torch.empty((plan.rows[level + 1] if pool else x.shape[0], 128),
                      dtype=torch.float32 if out_f32 else x.dtype, device=x.device)
    bd_h, b1_h = _host_f32(bd), _host_f32(b1)
    _lib.check(_lib.lib().mucon_wavenet_layer_bf16_ex(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(Wd_kco16), bd_h.ctypes.data_as(C.c_void_p), _lib.ptr(W1_kco16),
        b1_h.ctypes.data_as(C.c_void_p), _lib.ptr(tiles), C.c_int(n_tiles), C.c_int64(x.shape[0]),
        C.c_int64(out.shape[0]), C.c_int(dilation), C.c_int(int(pool)), C.c_int(int(relu_final)),
        C.c_int(int(out_f32)), C.c_int(int(x.dtype == torch.float16)), C.c_int(int(residual)), _stream(x.device)),
        "mucon_wavenet_layer_bf16_ex")
    return out


PADDED_WIDTHS = (32, 64)   # hidden sizes served by a zero-padded 128-channel twin (WaveNetBlock._padded_twin)
LAST_CONV_16 = os.environ.get("MUCON_LAST_CONV_16", "1") != "0"
DEAD_DILATION = 1 << 20   # no video is that long: both side taps only see padding and are never loaded


def conv1x1_bf16_rows(x, ident_k16, W_k16, bias_h, plan, level):
    """Plain 1x1 conv of NON-NEGATIVE 16-bit rows -> fp32 on the resident-weight layer pipeline: GEMM 1 multiplies with
    an identity centre tap (exact; relu(x) = x), GEMM 2 with W, no skip connection.  This is how the fast path runs
    last_conv (temporal.py:144-145; its input has just been through the ReLU of :144): 55 us instead of the 190 us the
    one-tap TF32 conv_gemm_kernel needs for the 2720 partly filled tiles of the c2 batch at T/16."""
    return wavenet_layer_bf16_rows(x, ident_k16, _ZERO_BIAS, W_k16, bias_h, plan, level, DEAD_DILATION, False, False,
                                   out_f32=True, residual=False)


def maxpool2_rows(x, plan, level, mode=0):
    """mode 0: max_pool1d(2); mode 1: avg_pool1d(2) * 2 (the sum of the pair, temporal.py:141-142)."""
    Cc = x.shape[1]
    out = torch.empty((plan.rows[level + 1], Cc), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().mucon_pool2(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(plan.off[level]), _lib.ptr(plan.off[level + 1]), C.c_int(plan.V),
        C.c_int(plan.max_T[level + 1]), C.c_int(Cc), C.c_int(int(mode)), _stream(x.device)), "mucon_pool2")
    return out


def groupnorm_relu_rows(x, gamma, beta, off, V, groups, eps, relu):
    out = torch.empty_like(x)
    _lib.check(_lib.lib().mucon_groupnorm_relu(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(off), C.c_int(V), C.c_int(x.shape[1]),
        C.c_int(groups), C.c_float(eps), C.c_int(int(relu)), _stream(x.device)), "mucon_groupnorm_relu")
    return out


def logsoftmax_expand_rows(logits, plan, z_level):
    Cc = logits.shape[1]
    out = torch.empty((plan.rows[0], Cc), dtype=torch.float32, device=logits.device)
    _lib.check(_lib.lib().mucon_logsoftmax_expand(
        _lib.ptr(logits), _lib.ptr(plan.off[z_level]), _lib.ptr(plan.off[0]), C.c_int(plan.V), C.c_int(plan.max_T[0]),
        C.c_int(Cc), _lib.ptr(out), _stream(logits.device)), "mucon_logsoftmax_expand")
    return out


def _tco(conv):
    """Conv1d weight [Cout, Cin, k] -> contiguous [k, Cin, Cout]."""
    return conv.weight.detach().permute(2, 1, 0).contiguous().float()


def _kco(conv):
    """Conv1d weight [Cout, Cin, k] -> contiguous [k*Cout, Cin] (tcgen05 B operand, K-major)."""
    w = conv.weight.detach().permute(2, 0, 1).contiguous().float()
    return w.view(w.shape[0] * w.shape[1], w.shape[2])


class WaveNetLayer(nn.Module):
    """Parameter container with the reference's names (temporal.py:9-53)."""

    def __init__(self, num_channels, kernel_size, dilation, drop=0.25, leaky=False):
        super().__init__()
        self.num_channels, self.kernel_size, self.dilation, self.leaky = num_channels, kernel_size, dilation, leaky
        self.dilated_conv = nn.Conv1d(num_channels, num_channels, kernel_size, dilation=dilation, padding=dilation)
        self.conv_1x1 = nn.Conv1d(num_channels, num_channels, kernel_size=1)
        self.drop = nn.Dropout(drop)


class WaveNetBlock(nn.Module):
    """Drop-in for core.modules.temporal.WaveNetBlock (temporal.py:77-147), forward on the GPU."""

    def __init__(self, in_channels, stages=(1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024), out_dims=64, kernel_size=3,
                 pooling=True, pooling_layers=(1, 2, 4, 8), pooling_type="max", dropout_rate=0.25, leaky=False):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("kernel_size != 3")
        # leaky ReLU (model.ft.leaky_relu) and the non-max pooling type (avg_pool1d * 2, temporal.py:141-142) run on
        # the TF32 conv GEMM / fp32 kernels with one launch per conv; the fused 16-bit layer kernels cover the
        # default configuration (ReLU, max pooling)
        self.in_channels, self.stages, self.out_dims = in_channels, list(stages), out_dims
        self.num_stages = len(self.stages)
        self.kernel_size, self.pooling, self.pooling_type = kernel_size, pooling, pooling_type
        self.pooling_layers, self.dropout_rate, self.leaky = list(pooling_layers), dropout_rate, leaky
        self.first_conv = nn.Conv1d(in_channels, out_dims, kernel_size=1)
        self.last_conv = nn.Conv1d(out_dims, out_dims, kernel_size=1)
        self.layers = []
        for i, d in enumerate(self.stages):
            layer = WaveNetLayer(out_dims, kernel_size, d, drop=dropout_rate, leaky=leaky)
            self.layers.append(layer)
            self.add_module("l_{}".format(i), layer)
        self._cache = None

    def n_pools(self):
        return sum(1 for i in range(self.num_stages) if self.pooling and i in self.pooling_layers)

    def _padded_twin(self):
        """WaveNetBlock(out_dims=128) holding this block's parameters zero-padded to 128 channels (cached by parameter
        versions; not a registered submodule, so state_dict() is unchanged)."""
        key = tuple(p._version for p in self.parameters()) + (str(self.first_conv.weight.device),)
        got = self.__dict__.get("_twin")
        if got is not None and got[0] == key:
            return got[1]
        h = self.out_dims
        twin = WaveNetBlock(self.in_channels, stages=self.stages, out_dims=128, kernel_size=self.kernel_size,
                            pooling=self.pooling, pooling_layers=self.pooling_layers, pooling_type=self.pooling_type,
                            dropout_rate=self.dropout_rate, leaky=self.leaky).to(self.first_conv.weight.device).eval()
        with torch.no_grad():
            for p in twin.parameters():
                p.zero_()
            twin.first_conv.weight[:h].copy_(self.first_conv.weight)
            twin.first_conv.bias[:h].copy_(self.first_conv.bias)
            for mine, theirs in zip(self.layers + [self], twin.layers + [twin]):
                convs = [("last_conv",)] if mine is self else [("dilated_conv",), ("conv_1x1",)]
                for (name,) in convs:
                    getattr(theirs, name).weight[:h, :h].copy_(getattr(mine, name).weight)
                    getattr(theirs, name).bias[:h].copy_(getattr(mine, name).bias)
        self.__dict__["_twin"] = (key, twin)
        return twin

    def _weights(self):
        key = tuple(p._version for p in self.parameters()) + (str(self.first_conv.weight.device),)
        if self._cache is None or self._cache[0] != key:
            w = dict(first_w=self.first_conv.weight.detach()[:, :, 0].contiguous().float(),
                     first_b=self.first_conv.bias.detach().contiguous().float(),
                     last_w=_tco(self.last_conv), last_b=self.last_conv.bias.detach().contiguous().float(),
                     layers=[(_tco(l.dilated_conv), l.dilated_conv.bias.detach().contiguous().float(),
                              _tco(l.conv_1x1), l.conv_1x1.bias.detach().contiguous().float()) for l in self.layers])
            if self.out_dims == 128:  # tensor-core layouts
                w["last_k"] = _kco(self.last_conv)
                w["layers_k"] = [(_kco(l.dilated_conv), _kco(l.conv_1x1)) for l in self.layers]
                w["layers_k16"] = [(a.to(torch.bfloat16).contiguous(), b.to(torch.bfloat16).contiguous())
                                   for a, b in w["layers_k"]]
                w["layers_kh16"] = [(a.to(torch.float16).contiguous(), b.to(torch.float16).contiguous())
                                    for a, b in w["layers_k"]]
                w["layers_bias_h"] = [(_host_f32(l.dilated_conv.bias), _host_f32(l.conv_1x1.bias)) for l in self.layers]
                # last_conv on the 16-bit layer pipeline (conv1x1_bf16_rows): an identity centre tap + its own weights
                ident = torch.zeros(3 * 128, 128, device=w["last_k"].device)
                ident[128:256] = torch.eye(128, device=ident.device)
                w["last_k16"] = {dt: (ident.to(dt).contiguous(), w["last_k"].to(dt).contiguous())
                                 for dt in (torch.float16, torch.bfloat16)}
                w["last_b_h"] = _host_f32(self.last_conv.bias)
            self._cache = (key, w)
        return self._cache[1]

    def project_packed(self, feats, precision=None):
        """first_conv + ReLU alone (temporal.py:133) in the activation type of the fused layer kernels; hand the result
        to forward_packed(..., x0=...).  Used by MuConBackbone.infer_pooled_pipelined."""
        precision = precision or DEFAULT_PRECISION
        if precision not in ("fp16", "bf16") or self.out_dims != 128 or self.in_channels % 32 != 0 or self.leaky:
            raise NotImplementedError("project_packed covers the default fast path (fp16 / bf16 layers, 128 channels)")
        w = self._weights()
        return gemm_tf32_bias_act(feats, w["first_w"], w["first_b"], relu=True,
                                  out_dtype=torch.float16 if precision == "fp16" else torch.bfloat16)

    def forward_packed(self, feats, plan, tensor_cores=True, fused_layers=True, precision=None, x0=None):
        """feats [sum T, in_channels] float32 rows (time-major, videos concatenated) -> [sum T', out_dims].
        x0: the projection's output if it has been computed already (project_packed).
        tensor_cores=False keeps the 128->128 convolutions on the fp32 FFMA kernels (exact fp32);
        precision: "fp16" | "bf16" | "tf32" (| "fp32" == tensor_cores=False), default DEFAULT_PRECISION."""
        if self.training and self.dropout_rate > 0:
            raise NotImplementedError("training-mode dropout is not implemented; call .eval()")
        if not feats.is_cuda:
            raise _lib.MuconError("the backbone needs CUDA tensors (there is no CPU fallback)")
        precision = precision or DEFAULT_PRECISION
        if precision not in ("fp16", "bf16", "tf32", "fp32"):
            raise ValueError(f"precision {precision!r}")
        if precision == "fp32":
            tensor_cores = False
        if tensor_cores and self.out_dims in PADDED_WIDTHS and self.in_channels % 32 == 0:
            # model.ft.hidden_size = 32 / 64 on the 128-channel tensor-core kernels: a 128-channel twin whose extra
            # channels have zero weights and biases computes exact zeros there (relu(0) = 0, 0 + 0 = 0, max(0, 0) = 0)
            # and never feeds them into the real ones
            z = self._padded_twin().forward_packed(feats, plan, tensor_cores=True, fused_layers=fused_layers,
                                                   precision=precision, x0=x0)
            return z[:, :self.out_dims].contiguous()
        act = 2 if self.leaky else 1                       # activation mode of the kernels' relu flags
        pool_mode = 0 if self.pooling_type == "max" else 1
        if act != 1 or pool_mode != 0:
            fused_layers = False
            if precision in ("fp16", "bf16"):
                precision = "tf32"
        w = self._weights()
        V = plan.V
        last = self.num_stages - 1
        if precision in ("bf16", "fp16") and tensor_cores and fused_layers and self.out_dims == 128 and self.num_stages > 0:
            dt = torch.float16 if precision == "fp16" else torch.bfloat16
            if x0 is not None:
                x = x0
            elif self.in_channels % 32 == 0:
                x = gemm_tf32_bias_act(feats, w["first_w"], w["first_b"], relu=True, out_dtype=dt)   # temporal.py:133
            else:
                x = conv1d_rows(feats, w["first_w"].t().contiguous()[None], w["first_b"], plan.off[0], V, plan.max_T[0],
                                relu_out=True).to(dt)
            level = 0
            for i, (wd, bd, w1, b1) in enumerate(w["layers"]):
                pooled = self.pooling and i in self.pooling_layers
                wdk, w1k = w["layers_kh16" if precision == "fp16" else "layers_k16"][i]
                bd, b1 = w["layers_bias_h"][i]
                # the ReLU of temporal.py:144 is folded into the last layer's store; without LAST_CONV_16 that layer
                # hands fp32 to the TF32 last_conv
                x = wavenet_layer_bf16_rows(x, wdk, bd, w1k, b1, plan, level, self.stages[i], pooled,
                                            relu_final=(i == last),
                                            out_f32=(i == last and not pooled and not LAST_CONV_16))
                if pooled:
                    level += 1
            if LAST_CONV_16 and x.dtype == dt:
                ident, wl = w["last_k16"][dt]
                return conv1x1_bf16_rows(x, ident, wl, w["last_b_h"], plan, level)                  # temporal.py:144-145
            if x.dtype != torch.float32:
                x = x.float()
            return conv_gemm_rows(x, w["last_k"], w["last_b"], plan, level)                          # temporal.py:144-145
        if self.out_dims == 128 and self.in_channels % 32 == 0:
            x = gemm_tf32_bias_act(feats, w["first_w"], w["first_b"], relu=act)        # temporal.py:133
        else:
            x = conv1d_rows(feats, w["first_w"].t().contiguous()[None], w["first_b"], plan.off[0], V, plan.max_T[0],
                            relu_out=act)
        level = 0
        tc = tensor_cores and self.out_dims == 128
        for i, (wd, bd, w1, b1) in enumerate(w["layers"]):
            off, mt = plan.off[level], plan.max_T[level]
            pooled = self.pooling and i in self.pooling_layers
            if tc and fused_layers:
                wdk, w1k = w["layers_k"][i]
                # dilated conv -> ReLU -> 1x1 -> + x (-> ReLU of temporal.py:144 after the last layer)
                # (-> max-pool) in one launch
                x = wavenet_layer_rows(x, wdk, bd, w1k, b1, plan, level, self.stages[i], pooled, relu_final=(i == last))
                if pooled:
                    level += 1
                continue
            if tc:
                wdk, w1k = w["layers_k"][i]
                y = conv_gemm_rows(x, wdk, bd, plan, level, dilation=self.stages[i], relu_mid=act)      # temporal.py:48-49
                # the ReLU in front of last_conv (temporal.py:144) is folded into the last layer's store
                x = conv_gemm_rows(y, w1k, b1, plan, level, residual=x, relu_final=act * (i == last and not pooled))
            else:
                y = conv1d_rows(x, wd, bd, off, V, mt, dilation=self.stages[i], relu_out=act)            # temporal.py:48-49
                x = conv1d_rows(y, w1, b1, off, V, mt, residual=x)                                        # temporal.py:50-52
            if pooled:
                x = maxpool2_rows(x, plan, level, pool_mode)                                               # temporal.py:139-142
                level += 1
        if tc:
            folded = fused_layers or not (self.pooling and last in self.pooling_layers)
            if not folded:
                x = torch.nn.functional.leaky_relu(x) if self.leaky else torch.relu(x)
            return conv_gemm_rows(x, w["last_k"], w["last_b"], plan, level)                               # temporal.py:144-145
        return conv1d_rows(x, w["last_w"], w["last_b"], plan.off[level], V, plan.max_T[level], relu_in=act)

    def forward(self, x):
        """x [B, in_channels, T] -> [B, out_dims, T']  (temporal.py:128-147)."""
        B, _, T = x.shape
        plan = BackbonePlan([T] * B, self.n_pools(), x.device)
        rows = x.detach().permute(0, 2, 1).reshape(B * T, self.in_channels).contiguous().float()
        z = self.forward_packed(rows, plan)
        Tz = plan.T[-1][0]
        return z.view(B, Tz, self.out_dims).permute(0, 2, 1).contiguous()


def conv_gemm_shifts_rows(x, W_kco, bias, shifts, plan, level, relu_mid=False, relu_final=False, residual=None):
    """128 -> 128 channel conv with explicit tap row-shifts on tcgen05.  W_kco: [len(shifts)*128, 128]."""
    out = torch.empty_like(x)
    sh = np.asarray(shifts, dtype=np.int32)
    _lib.check(_lib.lib().mucon_conv_gemm_tf32_shifts(
        _lib.ptr(x), _lib.ptr(out), _lib.ptr(W_kco), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(plan.tiles[level]),
        C.c_int(plan.n_tiles[level]), C.c_int64(x.shape[0]), sh.ctypes.data_as(C.c_void_p), C.c_int(int(sh.size)),
        C.c_int(int(relu_mid)), C.c_int(int(relu_final)), _stream(x.device)), "mucon_conv_gemm_tf32_shifts")
    return out


class NoFt(nn.Module):
    """Drop-in for core.modules.temporal.NoFt (temporal.py:56-74): a single 1x1 projection."""

    def __init__(self, in_chnnels, out_dims, kernel_size=1):
        super().__init__()
        if kernel_size != 1:
            raise NotImplementedError("NoFt with kernel_size != 1")
        self.in_chnnels, self.out_dims, self.kernel_size = in_chnnels, out_dims, kernel_size
        self.last_conv = nn.Conv1d(in_channels=in_chnnels, out_channels=out_dims, kernel_size=kernel_size)

    def n_pools(self):
        return 0

    def forward_packed(self, feats, plan, tensor_cores=True, fused_layers=True):
        if not feats.is_cuda:
            raise _lib.MuconError("the backbone needs CUDA tensors (there is no CPU fallback)")
        w = self.last_conv.weight.detach()[:, :, 0].contiguous().float()
        b = self.last_conv.bias.detach().contiguous().float()
        if tensor_cores and self.out_dims == 128 and self.in_chnnels % 32 == 0:
            return gemm_tf32_bias_act(feats, w, b, relu=False)
        return conv1d_rows(feats, w.t().contiguous()[None], b, plan.off[0], plan.V, plan.max_T[0])

    def forward(self, x):
        B, _, T = x.shape
        plan = BackbonePlan([T] * B, 0, x.device)
        rows = x.detach().permute(0, 2, 1).reshape(B * T, self.in_chnnels).contiguous().float()
        return self.forward_packed(rows, plan).view(B, T, self.out_dims).permute(0, 2, 1).contiguous()


class MSTCNPPFirstStage(nn.Module):
    """Drop-in for core.modules.temporal.MSTCNPPFirstStage (temporal.py:150-204), forward on the GPU.

    Layer i is  f <- relu(conv_fusion(cat(conv_dilated_1(f), conv_dilated_2(f)))) + f  with dilations
    2**(L-1-i) and 2**i.  Everything in front of the ReLU is linear, so the three convolutions fold (on
    the host, in float64) into ONE convolution with taps at {-d1, -d2, 0, +d2, +d1}:
    W_eff[shift] = Wf[:, :H] @ W1[tap] (+ Wf[:, H:] @ W2[tap] where shifts coincide), and the layer is a
    single tcgen05 launch (mucon_conv_gemm_tf32_shifts: bias + ReLU + residual fused).  Rounding differs
    from the reference's two-stage evaluation by less than the TF32 operand rounding."""

    def __init__(self, num_layers, num_f_maps, input_dim, output_dim, pooling_layers=(1, 2, 4, 8)):
        super().__init__()
        if num_f_maps != 128 or output_dim != 128:
            raise NotImplementedError("the tcgen05 conv kernels are built for 128 feature maps")
        self.num_layers, self.num_f_maps, self.input_dim, self.output_dim = num_layers, num_f_maps, input_dim, output_dim
        self.conv_1x1_in = nn.Conv1d(input_dim, num_f_maps, 1)
        self.conv_dilated_1 = nn.ModuleList(
            nn.Conv1d(num_f_maps, num_f_maps, 3, padding=2 ** (num_layers - 1 - i), dilation=2 ** (num_layers - 1 - i))
            for i in range(num_layers))
        self.conv_dilated_2 = nn.ModuleList(
            nn.Conv1d(num_f_maps, num_f_maps, 3, padding=2 ** i, dilation=2 ** i) for i in range(num_layers))
        self.conv_fusion = nn.ModuleList(nn.Conv1d(2 * num_f_maps, num_f_maps, 1) for i in range(num_layers))
        self.dropout = nn.Dropout()
        self.conv_out = nn.Conv1d(num_f_maps, output_dim, 1)
        self.pooling_layers = list(pooling_layers)
        self._cache = None

    def n_pools(self):
        return sum(1 for i in range(self.num_layers) if i in self.pooling_layers)

    def _weights(self):
        key = tuple(p._version for p in self.parameters()) + (str(self.conv_out.weight.device),)
        if self._cache is None or self._cache[0] != key:
            H = self.num_f_maps
            layers = []
            for i in range(self.num_layers):
                d1, d2 = 2 ** (self.num_layers - 1 - i), 2 ** i
                W1 = self.conv_dilated_1[i].weight.detach().double()   # [Co, Ci, 3]
                W2 = self.conv_dilated_2[i].weight.detach().double()
                Wf = self.conv_fusion[i].weight.detach().double()[:, :, 0]
                Wf1, Wf2 = Wf[:, :H], Wf[:, H:]
                taps = {}
                for t in range(3):
                    taps[(t - 1) * d1] = taps.get((t - 1) * d1, 0) + Wf1 @ W1[:, :, t]
                    taps[(t - 1) * d2] = taps.get((t - 1) * d2, 0) + Wf2 @ W2[:, :, t]
                shifts = sorted(taps)
                W = torch.cat([taps[sft] for sft in shifts], 0).float().contiguous()   # [n*Co, Ci]
                bias = (Wf1 @ self.conv_dilated_1[i].bias.detach().double()
                        + Wf2 @ self.conv_dilated_2[i].bias.detach().double()
                        + self.conv_fusion[i].bias.detach().double()).float().contiguous()
                layers.append((shifts, W, bias))
            w = dict(in_w=self.conv_1x1_in.weight.detach()[:, :, 0].contiguous().float(),
                     in_b=self.conv_1x1_in.bias.detach().contiguous().float(),
                     out_k=_kco(self.conv_out), out_b=self.conv_out.bias.detach().contiguous().float(), layers=layers)
            self._cache = (key, w)
        return self._cache[1]

    def forward_packed(self, feats, plan, tensor_cores=True, fused_layers=True):
        """feats [sum T, input_dim] float32 rows -> [sum T', output_dim]."""
        if self.training:
            raise NotImplementedError("training-mode dropout is not implemented; call .eval()")
        if not feats.is_cuda:
            raise _lib.MuconError("the backbone needs CUDA tensors (there is no CPU fallback)")
        if not tensor_cores:
            raise NotImplementedError("MSTCNPPFirstStage runs on the tcgen05 (TF32) kernels only")
        w = self._weights()
        if self.input_dim % 32 == 0:
            f = gemm_tf32_bias_act(feats, w["in_w"], w["in_b"], relu=False)                   # temporal.py:187
        else:
            f = conv1d_rows(feats, w["in_w"].t().contiguous()[None], w["in_b"], plan.off[0], plan.V, plan.max_T[0])
        level = 0
        for i, (shifts, W, bias) in enumerate(w["layers"]):
            f = conv_gemm_shifts_rows(f, W, bias, shifts, plan, level, relu_mid=True, residual=f)   # :189-196
            if i in self.pooling_layers:
                f = maxpool2_rows(f, plan, level)                                               # :198-199
                level += 1
        return conv_gemm_rows(f, w["out_k"], w["out_b"], plan, level)                          # :201

    def forward(self, x):
        B, _, T = x.shape
        plan = BackbonePlan([T] * B, self.n_pools(), x.device)
        rows = x.detach().permute(0, 2, 1).reshape(B * T, self.input_dim).contiguous().float()
        z = self.forward_packed(rows, plan)
        return z.view(B, int(plan.T[-1][0]), self.output_dim).permute(0, 2, 1).contiguous()


class MuConBackbone(nn.Module):
    """The backbone-side members of the reference model, under the reference's attribute names:
    `ft` (models.py:162-171), `ft_last_gn` (:188-191), `conv_classifier` (:276-278)."""

    def __init__(self, input_feature_size=2048, num_classes=48, hidden_size=128,
                 stages=(1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024), pooling=True, pooling_layers=(1, 2, 4, 8),
                 last_gn=True, last_gn_num_groups=32, last_relu=True, ft_type="wavenet", leaky_relu=False,
                 pooling_type="max", dropout_rate=0.25):
        super().__init__()
        self.num_classes, self.hidden_size = num_classes, hidden_size
        self.last_gn, self.last_relu = last_gn, last_relu
        if ft_type == "wavenet":      # models.py:160-186
            self.ft = WaveNetBlock(input_feature_size, stages=stages, out_dims=hidden_size, pooling=pooling,
                                   pooling_layers=pooling_layers, pooling_type=pooling_type, leaky=leaky_relu,
                                   dropout_rate=dropout_rate)
        elif ft_type == "mstcnpp":
            self.ft = MSTCNPPFirstStage(input_dim=input_feature_size, num_layers=len(stages), output_dim=hidden_size,
                                        num_f_maps=hidden_size, pooling_layers=pooling_layers)
        elif ft_type == "noft":
            self.ft = NoFt(in_chnnels=input_feature_size, out_dims=hidden_size)
        else:
            raise Exception(f"Invalid ft type ({ft_type})")
        self.ft_last_gn = nn.GroupNorm(num_groups=last_gn_num_groups, num_channels=hidden_size)
        self.conv_classifier = nn.Conv1d(hidden_size, num_classes, kernel_size=1)

    def plan(self, T, device=None):
        return BackbonePlan(T, self.ft.n_pools(), device or self.conv_classifier.weight.device)

    # ---- packed (variable-length batch) API ----------------------------------------------------
    def encode_packed(self, feats, plan, tensor_cores=True, fused_layers=True, precision=None):
        """temporal_modeling_forward for a packed batch: [sum T, D] -> [sum Tz, hidden]."""
        kw = dict(precision=precision) if isinstance(self.ft, WaveNetBlock) else {}
        z = self.ft.forward_packed(feats, plan, tensor_cores=tensor_cores, fused_layers=fused_layers, **kw)
        lvl = len(plan.off) - 1
        if self.last_gn:
            z = groupnorm_relu_rows(z, self.ft_last_gn.weight.detach().float(), self.ft_last_gn.bias.detach().float(),
                                    plan.off[lvl], plan.V, self.ft_last_gn.num_groups, self.ft_last_gn.eps,
                                    relu=self.last_relu)
        elif self.last_relu:
            z = torch.relu(z)
        return z

    def logprobs_packed(self, z, plan):
        """frame_classifier_forward + predict's log_softmax: [sum Tz, hidden] -> [sum T, classes].  The rows are those
        of logprobs_pooled_packed (same kernel arithmetic), written for every frame through the nearest-neighbour
        index of F.interpolate."""
        lvl = len(plan.off) - 1
        w = self.conv_classifier.weight.detach().permute(2, 1, 0).contiguous().float()
        logits = conv1d_rows(z, w, self.conv_classifier.bias.detach().float(), plan.off[lvl], plan.V, plan.max_T[lvl])
        return logsoftmax_expand_rows(logits, plan, lvl)

    def infer_pooled_pipelined(self, feats, T, n_chunks=4, proj_sms=88, precision=None):
        """infer_pooled_packed with the batch cut into n_chunks contiguous groups of videos and two streams: the
        projection of chunk k+1 (HBM-bound; it needs about two thirds of the SMs to pull its stream through the L2) runs
        SIDE BY SIDE with the layer / tail kernels of chunk k (tensor-bound) on disjoint SMs -- every kernel here is
        persistent with one CTA per SM, so the grids are capped (mucon_set_sm_limit) at proj_sms and SMs - proj_sms.
        Returns (table [sum Tz, classes], row offsets [V+1]) like infer_pooled_packed (same values)."""
        dev = feats.device
        T = np.asarray(T, dtype=np.int64)
        key = (tuple(T.tolist()), n_chunks, str(dev))
        pc = getattr(self, "_pipe_cache", None)
        if pc is None or pc[0] != key:
            cum = np.cumsum(T)
            cuts = [0] + [int(np.searchsorted(cum, cum[-1] * (k + 1) / n_chunks, side="left")) + 1 for k in range(n_chunks - 1)]
            cuts = sorted(set(min(max(c, 0), len(T)) for c in cuts)) + [len(T)]
            cuts = [c for i, c in enumerate(cuts) if i == 0 or c > cuts[i - 1]]
            plans = [self.plan(T[a:b], dev) for a, b in zip(cuts[:-1], cuts[1:])]
            rows = np.concatenate([[0], np.cumsum([int(T[a:b].sum()) for a, b in zip(cuts[:-1], cuts[1:])])])
            full = self.plan(T, dev)
            pc = self._pipe_cache = (key, plans, rows, full, torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        _, plans, rows, full, s_proj, s_lay = pc
        lib = _lib.lib()
        sms = int(lib.mucon_device_sm_count())
        main = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(main)
        s_proj.wait_event(start)
        s_lay.wait_event(start)
        tables = []
        try:
            for k, pl in enumerate(plans):
                with torch.cuda.stream(s_proj):
                    lib.mucon_set_sm_limit(C.c_int(proj_sms))
                    x0 = self.ft.project_packed(feats[int(rows[k]):int(rows[k + 1])], precision)
                    ev = torch.cuda.Event()
                    ev.record(s_proj)
                x0.record_stream(s_lay)
                with torch.cuda.stream(s_lay):
                    s_lay.wait_event(ev)
                    lib.mucon_set_sm_limit(C.c_int(max(sms - proj_sms, 8)))
                    tb, _ = self.infer_pooled_packed(None, pl, precision=precision, x0=x0)
                    tb.record_stream(main)
                tables.append(tb)
        finally:
            lib.mucon_set_sm_limit(C.c_int(0))   # never leave the cap behind: every later launch of this thread would obey it
        done = torch.cuda.Event()
        done.record(s_lay)
        main.wait_event(done)
        return torch.cat(tables), full.off[len(full.off) - 1]

    def infer_pooled_packed(self, feats, plan, precision=None, want_z=False, x0=None):
        """Features -> pooled-resolution log-probabilities in as few launches as the path has: ft (projection + one
        launch per layer + last_conv), then GroupNorm statistics and ONE launch for GroupNorm + ReLU + classifier +
        log_softmax (mucon_tail_logprobs).  Returns (table [sum Tz, classes], row offsets [V+1]) (+ z if want_z)."""
        lvl = len(plan.off) - 1
        if not (self.last_gn and self.hidden_size == 128 and self.num_classes <= 64):
            z = self.encode_packed(feats, plan, precision=precision)
            table, off = self.logprobs_pooled_packed(z, plan)
            return (table, off, z) if want_z else (table, off)
        kw = dict(precision=precision) if isinstance(self.ft, WaveNetBlock) else {}
        if x0 is not None:
            kw["x0"] = x0
        x = self.ft.forward_packed(feats if feats is not None else x0, plan, **kw)
        key = (self.conv_classifier.weight._version, self.conv_classifier.bias._version, str(x.device))
        if getattr(self, "_cls_cache", None) is None or self._cls_cache[0] != key:
            self._cls_cache = (key, self.conv_classifier.weight.detach()[:, :, 0].t().contiguous().float(),
                               self.conv_classifier.bias.detach().contiguous().float())
        _, wc, bc = self._cls_cache
        table = torch.empty((x.shape[0], self.num_classes), dtype=torch.float32, device=x.device)
        z = torch.empty_like(x) if want_z else None
        stats = torch.empty(2 * plan.V * self.ft_last_gn.num_groups, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mucon_tail_logprobs(
            _lib.ptr(x), _lib.ptr(plan.off[lvl]), _lib.ptr(plan.tiles[lvl]), _lib.ptr(plan.tile_vid[lvl]),
            C.c_int(plan.n_tiles[lvl]), C.c_int(plan.V), C.c_int(self.hidden_size), C.c_int(self.ft_last_gn.num_groups),
            C.c_float(self.ft_last_gn.eps), C.c_int(int(self.last_relu)), _lib.ptr(self.ft_last_gn.weight.detach()),
            _lib.ptr(self.ft_last_gn.bias.detach()), _lib.ptr(wc), _lib.ptr(bc), C.c_int(self.num_classes),
            _lib.ptr(stats), _lib.ptr(z), _lib.ptr(table), _stream(x.device)), "mucon_tail_logprobs")
        return (table, plan.off[lvl], z) if want_z else (table, plan.off[lvl])

    def logprobs_pooled_packed(self, z, plan):
        """The same log-probabilities at the POOLED resolution, not expanded: [sum Tz, hidden] -> ([sum Tz, classes],
        row offsets [V+1]).  ViterbiEngine.run(..., z_off=offsets) aligns straight from this table (bit-identical
        to the expanded path; the [sum T, classes] array is never written or read)."""
        lvl = len(plan.off) - 1
        w = self.conv_classifier.weight.detach().permute(2, 1, 0).contiguous().float()
        logits = conv1d_rows(z, w, self.conv_classifier.bias.detach().float(), plan.off[lvl], plan.V, plan.max_T[lvl])
        out = torch.empty_like(logits)
        _lib.check(_lib.lib().mucon_logsoftmax_rows(
            _lib.ptr(logits), C.c_int64(logits.shape[0]), C.c_int(logits.shape[1]), _lib.ptr(out),
            _stream(logits.device)), "mucon_logsoftmax_rows")
        return out, plan.off[lvl]

    # ---- the reference's signatures (batch size 1) -----------------------------------------------
    def temporal_modeling_forward(self, input):
        """[B, T, D] -> [B, T', D']  (models.py:746-773), eval mode."""
        B, T, D = input.shape
        plan = self.plan([T] * B, input.device)
        z = self.encode_packed(input.detach().reshape(B * T, D).contiguous().float(), plan)
        return z.view(B, -1, self.hidden_size)

    def frame_classifier_forward(self, temporal_encoded, target_length):
        """[1, Ds, Tz] -> [1, num_classes, Tf] logits (models.py:567-582)."""
        z = temporal_encoded.detach()[0].t().contiguous().float()  # [Tz, Ds]
        Tz = z.shape[0]
        off = torch.tensor([0, Tz], dtype=torch.int64, device=z.device)
        w = self.conv_classifier.weight.detach().permute(2, 1, 0).contiguous().float()
        logits = conv1d_rows(z, w, self.conv_classifier.bias.detach().float(), off, 1, Tz)
        idx = torch.clamp(torch.floor(torch.arange(target_length, device=z.device, dtype=torch.float32)
                                      * (float(np.float32(Tz) / np.float32(target_length)))).long(), max=Tz - 1)
        return logits[idx].t().unsqueeze(0).contiguous()
