"""The mutual-consistency loss around the GPU mask kernel.

Mirrors reference src/mucon/models.py:414-450 (`MuCon.mucon_loss`) and :452-525
(`calculate_mucon_loss_using_masks`, "flint" and "arithmetic" variants): softmax-projected lengths ->
`create_masks` (CUDA, autograd-aware, scales its argument in place like the reference) -> masked
class evidence -> NLL.  Everything after the masks is the same handful of torch ops the reference
uses, so gradients w.r.t. both the lengths and the frame logits flow as they do there.
"""
import torch
import torch.nn.functional as F

from .masks import create_masks, project_lengths_softmax


def loss_from_masks(absolute_lengths, masks, segmentation, target_transcript, mucon_type="flint", class_weight=None):
    """models.py:452-525.  masks [N, T], segmentation [T, C] logits, target_transcript [N] long."""
    if mucon_type == "flint":
        # p_i = log_softmax((mask_i . seg) / L_i)  (models.py:459-468), all segments at once
        evidence = (masks @ segmentation) / absolute_lengths[:, None]
        return F.nll_loss(F.log_softmax(evidence, dim=1), target_transcript, weight=class_weight, reduction="mean")
    if mucon_type == "arithmetic":
        # sum_i sum_t CE(seg_t, tr_i) * mask_i[t] / T  (models.py:489-523)
        logp = F.log_softmax(segmentation, dim=1)
        ce = -logp[:, target_transcript]  # [T, N]
        if class_weight is not None:
            ce = ce * class_weight[target_transcript][None, :]
        return (ce * masks.t()).sum() / segmentation.size(0)
    raise Exception(f"Invalid mucon type ({mucon_type})")


def mucon_loss(lengths, segmentation, target_transcript, template="box", overlap=0.0, mucon_type="flint",
               class_weight=None, align_corners=None):
    """models.py:414-450 given the s-head length logits [N], the frame logits [T, C] and the target
    transcript [N]."""
    T = segmentation.shape[0]
    absolute_lengths = project_lengths_softmax(T=T, L=lengths)
    masks = create_masks(T=T, L=absolute_lengths, template=template, overlap=overlap, align_corners=align_corners)
    return loss_from_masks(absolute_lengths, masks, segmentation, target_transcript, mucon_type, class_weight)
