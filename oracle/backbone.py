"""Torch fp32 functional restatement of the backbone forward -- ORACLE / TEST INFRASTRUCTURE ONLY.

Follows reference src/core/modules/temporal.py:43-53,128-147 (WaveNetLayer / WaveNetBlock.forward),
src/mucon/models.py:746-773 (GroupNorm + ReLU tail), :567-582 (nearest interpolate + 1x1 classifier)
and :368 (log_softmax), in eval mode (dropout off).  Works on a state_dict with the reference's
parameter names, so the same weights drive the reference module, this oracle and the CUDA path.
"""
import torch
import torch.nn.functional as F


def wavenet_block(sd, x, stages, pooling_layers, prefix="ft."):
    """x [B, Cin, T] -> [B, H, T']"""
    x = F.relu(F.conv1d(x, sd[prefix + "first_conv.weight"], sd[prefix + "first_conv.bias"]))
    for i, d in enumerate(stages):
        p = f"{prefix}l_{i}."
        y = F.conv1d(x, sd[p + "dilated_conv.weight"], sd[p + "dilated_conv.bias"], dilation=d, padding=d)
        y = F.relu(y)
        y = F.conv1d(y, sd[p + "conv_1x1.weight"], sd[p + "conv_1x1.bias"])
        x = y + x
        if i in pooling_layers:
            x = F.max_pool1d(x, kernel_size=2)
    x = F.relu(x)
    return F.conv1d(x, sd[prefix + "last_conv.weight"], sd[prefix + "last_conv.bias"])


def encode(sd, feats, stages, pooling_layers, groups=32, eps=1e-5):
    """temporal_modeling_forward: feats [1, T, D] -> [1, Tz, H]"""
    z = wavenet_block(sd, feats.permute(0, 2, 1), stages, pooling_layers)
    z = F.group_norm(z, groups, sd["ft_last_gn.weight"], sd["ft_last_gn.bias"], eps)
    z = F.relu(z)
    return z.permute(0, 2, 1)


def logprobs(sd, z, T):
    """frame_classifier_forward + log_softmax: z [1, Tz, H] -> [T, C]"""
    up = F.interpolate(z.permute(0, 2, 1), T)
    seg = F.conv1d(up, sd["conv_classifier.weight"], sd["conv_classifier.bias"])
    return F.log_softmax(seg.squeeze(0).permute(1, 0), dim=1)
