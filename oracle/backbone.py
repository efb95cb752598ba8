"""Torch fp32 functional restatement of the backbone forward -- ORACLE / TEST INFRASTRUCTURE ONLY.

Follows reference src/core/modules/temporal.py:43-53,128-147 (WaveNetLayer / WaveNetBlock.forward),
src/mucon/models.py:746-773 (GroupNorm + ReLU tail), :567-582 (nearest interpolate + 1x1 classifier)
and :368 (log_softmax), in eval mode (dropout off).  Works on a state_dict with the reference's
parameter names, so the same weights drive the reference module, this oracle and the CUDA path.
"""
import torch
import torch.nn.functional as F


def tf32_trunc(t):
    """what tcgen05.mma.kind::tf32 reads of an fp32 operand: the upper 19 bits (the 13 low mantissa bits are dropped)"""
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


class _Tf32Conv(torch.autograd.Function):
    """conv1d whose three GEMMs (forward, data gradient, weight gradient) see TF32-truncated operands and accumulate
    in fp32: the arithmetic of the CUDA training path, so that its kernels can be checked to accumulation-order
    accuracy instead of to TF32 accuracy."""

    @staticmethod
    def forward(ctx, x, w, b, dil, pad):
        ctx.save_for_backward(x, w)
        ctx.dil, ctx.pad = dil, pad
        return F.conv1d(tf32_trunc(x), tf32_trunc(w), b, dilation=dil, padding=pad)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.nn.grad.conv1d_input(x.shape, tf32_trunc(w), tf32_trunc(dy), dilation=ctx.dil, padding=ctx.pad)
        dw = torch.nn.grad.conv1d_weight(tf32_trunc(x), w.shape, tf32_trunc(dy), dilation=ctx.dil, padding=ctx.pad)
        return dx, dw, dy.sum((0, 2)), None, None


def wavenet_block(sd, x, stages, pooling_layers, prefix="ft.", tf32=False):
    """x [B, Cin, T] -> [B, H, T'].  tf32=True: every convolution through _Tf32Conv."""
    conv = (lambda x, w, b, d=1, p=0: _Tf32Conv.apply(x, w, b, d, p)) if tf32 else \
        (lambda x, w, b, d=1, p=0: F.conv1d(x, w, b, dilation=d, padding=p))
    x = F.relu(conv(x, sd[prefix + "first_conv.weight"], sd[prefix + "first_conv.bias"]))
    for i, d in enumerate(stages):
        p = f"{prefix}l_{i}."
        y = conv(x, sd[p + "dilated_conv.weight"], sd[p + "dilated_conv.bias"], d, d)
        y = F.relu(y)
        y = conv(y, sd[p + "conv_1x1.weight"], sd[p + "conv_1x1.bias"])
        x = y + x
        if i in pooling_layers:
            x = F.max_pool1d(x, kernel_size=2)
    x = F.relu(x)
    return conv(x, sd[prefix + "last_conv.weight"], sd[prefix + "last_conv.bias"])


def encode(sd, feats, stages, pooling_layers, groups=32, eps=1e-5, tf32=False):
    """temporal_modeling_forward: feats [1, T, D] -> [1, Tz, H]"""
    z = wavenet_block(sd, feats.permute(0, 2, 1), stages, pooling_layers, tf32=tf32)
    z = F.group_norm(z, groups, sd["ft_last_gn.weight"], sd["ft_last_gn.bias"], eps)
    z = F.relu(z)
    return z.permute(0, 2, 1)


def logprobs(sd, z, T):
    """frame_classifier_forward + log_softmax: z [1, Tz, H] -> [T, C]"""
    up = F.interpolate(z.permute(0, 2, 1), T)
    seg = F.conv1d(up, sd["conv_classifier.weight"], sd["conv_classifier.bias"])
    return F.log_softmax(seg.squeeze(0).permute(1, 0), dim=1)
