"""Soft-mask CT embedder (rows a10 of SURVEY.md section 8): produces the per-token weight ``w`` and ``y2``.

Interface and parameter names follow the reference's ``CT_Encoder`` (block/CT_encoder.py:5-44)
and ``VisionEmbedding`` (block/visionEmbedding.py:3-72) so the shipped
``pretrain_ct_vision_embedder/*.pt`` state dicts load unchanged:
``vision_embedding.proj.{weight,bias}``, ``vision_embedding.mask_token``, ``fc.0/2.{weight,bias}``,
``norm.{weight,bias}``.  It runs once per batch (not per timestep), ~16 k parameters, so it
stays plain PyTorch; the channel max/mean pooling is written as reductions instead of the
reference's ``AdaptiveMax/AvgPool2d((T,1))`` modules (same values).
"""
from __future__ import annotations

import torch
from torch import nn


class VisionEmbedding(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, contain_mask_token=False,
                 prepend_cls_token=False):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.patch_shape = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if contain_mask_token else None
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if prepend_cls_token else None

    def num_position_embeddings(self):
        return self.num_patches + (0 if self.cls_token is None else 1)

    def forward(self, x, masked_position=None, **kwargs):
        B, C, H, W = x.shape
        if (H, W) != self.img_size:
            raise AssertionError(f"Input image size ({H}*{W}) doesn't match model "
                                 f"({self.img_size[0]}*{self.img_size[1]}).")
        x = self.proj(x).flatten(2).transpose(1, 2)
        if masked_position is not None:
            assert self.mask_token is not None
            m = masked_position.unsqueeze(-1).type_as(self.mask_token)
            x = x * (1 - m) + self.mask_token.expand(B, x.shape[1], -1) * m
        if self.cls_token is not None:
            x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        return x


class CT_Encoder(nn.Module):
    def __init__(self, img_size=28, patch_size=2, in_channels=4, embed_dim=1024, contain_mask_token=True,
                 reduction_ratio=14):
        super().__init__()
        self.vision_embedding = VisionEmbedding(img_size=img_size, patch_size=patch_size, in_chans=in_channels,
                                                embed_dim=embed_dim, contain_mask_token=contain_mask_token)
        T = int((img_size / patch_size) ** 2)
        self.fc = nn.Sequential(nn.Linear(T, int(T / reduction_ratio)), nn.ReLU(inplace=True),
                                nn.Linear(int(T / reduction_ratio), T))
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x: torch.Tensor):
        """x (N,4,H,W) -> (weight (N,T,1) in (0,1), y2 (N,T,embed_dim))."""
        x = self.vision_embedding(x)                       # (N,T,D)
        max_out = self.fc(x.amax(dim=-1))                  # pooled over channels, MLP over the token axis
        avg_out = self.fc(x.mean(dim=-1))
        weight = torch.sigmoid(avg_out + max_out).unsqueeze(-1)
        return weight, self.norm(x * weight)
