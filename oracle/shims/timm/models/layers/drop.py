"""Stochastic depth as timm 0.4.12 ``layers/drop.py`` publishes it."""
import torch
import torch.nn as nn


class DropPath(nn.Module):
    """Per-sample residual-branch drop: ``x / keep * floor(keep + U[0,1))``."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.training or not self.drop_prob:
            return x
        keep = 1.0 - self.drop_prob
        mask_shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        gate = torch.rand(mask_shape, dtype=x.dtype, device=x.device).add_(keep).floor_()
        return x.div(keep) * gate
