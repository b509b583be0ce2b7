"""Drop-in for the reference's `code/loss.py` (imported by code/run_train_erc.py:16)."""
import _bootstrap  # noqa: F401
import torch
import torch.nn as nn
from mmdfn_b200.modules import FocalLoss  # noqa: F401


class MaskedNLLLoss(nn.Module):
    """code/loss.py:38-58 -- used only by the non-graph baselines; kept so the trainer's import succeeds."""

    def __init__(self, weight=None):
        super().__init__()
        self.weight = weight
        self.loss = nn.NLLLoss(weight=weight, reduction='sum')

    def forward(self, pred, target, mask):
        mask_ = mask.view(-1, 1)
        if self.weight is None:
            return self.loss(pred * mask_, target) / torch.sum(mask)
        return self.loss(pred * mask_, target) / torch.sum(self.weight[target] * mask_.squeeze())
