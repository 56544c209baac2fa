"""Micro stand-in for timm.models.vision_transformer (model.py:6, mamba_block.py:4): ViT Attention and Mlp
with timm 1.0.3 parameter names (qkv, proj / fc1, fc2), enough for the DiT baseline block."""
import torch
import torch.nn.functional as F


class Attention(torch.nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, **_):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = torch.nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = torch.nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        x = F.scaled_dot_product_attention(q, k, v)
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class Mlp(torch.nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=torch.nn.GELU, drop=0.0, **_):
        super().__init__()
        self.fc1 = torch.nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = torch.nn.Linear(hidden_features or in_features, out_features or in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))
