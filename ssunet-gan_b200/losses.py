"""Segmentation losses on the hot path (reference: losses.py:130-136, 274-302)."""
import torch
from torch import nn

from . import ops

__all__ = ["BCEDiceLoss", "StableBCELoss"]


class StableBCELoss(nn.Module):
    """mean(max(x,0) - x*t + log(1 + exp(-|x|)))   (losses.py:130-136)."""

    def forward(self, input, target):
        return ops.seg_losses(input, target)[2]


class BCEDiceLoss(nn.Module):
    """0.5 * StableBCE + (1 - mean_n dice_n), falling back to 2 * dice when the BCE term is NaN/Inf
    (losses.py:274-302).  One streaming kernel produces every partial sum; the NaN/Inf branch is a
    device-side select, so there is no host sync."""

    def forward(self, input, target):
        return ops.seg_losses(input, target)[0]


class BCEDiceAndContentLoss(nn.Module):
    """(BCEDiceLoss, nn.MSELoss) of the same logits in one pass (train_seg_gan.py:194-195)."""

    def forward(self, input, target):
        out = ops.seg_losses(input, target)
        return out[0], out[1]


class BCEWithLogitsConst(nn.Module):
    """nn.BCEWithLogitsLoss()(x, full_like(x, value))  (train_seg_gan.py:204,221-222)."""

    def forward(self, input, value):
        return ops.bce_with_logits_const(input, value)
