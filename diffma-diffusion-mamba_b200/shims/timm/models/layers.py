"""Micro stand-in for timm.models.layers (model.py:7): only to_2tuple and DropPath are imported."""
import torch


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(torch.nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert self.drop_prob == 0.0 or not self.training
        return x
