"""Mutual-consistency loss -- ORACLE / TEST INFRASTRUCTURE ONLY.  Restates reference
src/mucon/models.py:414-525 with the reference's own loop structure, on top of oracle.masks."""
import torch
import torch.nn.functional as F

from . import masks as omasks


def mucon_loss(lengths, segmentation, target_transcript, template="box", overlap=0.0, mucon_type="flint"):
    T = segmentation.shape[0]
    absolute = omasks.project_lengths_softmax(T, lengths)
    masks, scaled = omasks.create_masks_torch(T, absolute, overlap, template, align_corners=False)
    absolute = scaled  # the reference's create_masks scales its argument in place (masks.py:61)
    N = absolute.shape[0]
    if mucon_type == "flint":
        preds = []
        for i in range(N):
            window = (masks[i].unsqueeze(1) * segmentation).sum(0) / absolute[i]
            preds.append(F.log_softmax(window, dim=0))
        return F.nll_loss(torch.stack(preds), target_transcript, reduction="mean")
    losses = 0
    for i in range(N):
        target = target_transcript[i].repeat(T).long()
        losses = losses + (F.cross_entropy(segmentation, target, reduction="none") * masks[i]).sum()
    return losses / T
