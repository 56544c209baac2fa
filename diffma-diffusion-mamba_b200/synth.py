"""Deterministic synthetic weights and inputs shared by tests, bench and goldens.

There is no network, hence no checkpoints: benchmarks and parity tests use "trained-like" weights
generated from the parameter NAME (so the reference model and this package's mirror get identical
values as long as their state-dict keys agree, which is itself a drop-in requirement).  The
reference's own init is useless for this: it zeroes every adaLN / final layer so each block is the
identity.
"""
from __future__ import annotations

import math
import zlib

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


@torch.no_grad()
def fill_trained_like_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Overwrite every parameter with name-keyed pseudo-random values of a plausible trained scale."""
    for name, p in module.named_parameters():
        if name == "pos_embed":
            continue                                   # frozen sin-cos table, part of the architecture
        g = _gen(name, seed)
        leaf = name.rsplit(".", 1)[-1]
        shape = tuple(p.shape)
        if leaf == "A_log":
            if p.dim() == 2:                           # Mamba-1: S4D-real + jitter so A_n != -(n+1)
                base = torch.log(torch.arange(1, shape[1] + 1, dtype=torch.float32)).expand(shape)
                val = base + 0.3 * torch.randn(shape, generator=g)
            else:                                      # Mamba-2: one A per head, U(1,16)
                val = torch.log(1.0 + 15.0 * torch.rand(shape, generator=g))
        elif leaf == "D":
            val = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "dt_bias" or name.endswith("dt_proj.bias"):
            dt = torch.exp(torch.rand(shape, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
            val = dt + torch.log(-torch.expm1(-dt))    # inverse softplus: softplus(val) in [1e-3, 0.1]
        elif leaf == "bias":
            val = 0.02 * torch.randn(shape, generator=g)
        elif p.dim() == 1:
            val = 1.0 + 0.1 * torch.randn(shape, generator=g)     # LayerNorm / RMSNorm gains
        elif name.endswith("attention_network.3.weight"):
            val = torch.randn(shape, generator=g) * 0.05
        else:
            fan_in = p[0].numel()
            val = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        p.copy_(val.to(p.dtype))
    return module


def synthetic_batch(batch: int, input_size: int = 28, hidden: int = 512, tokens: int = 196, seed: int = 0,
                    device="cpu", dtype=torch.float32):
    """Latents / timesteps / conditioning of the shapes ``DiffMa.forward`` takes (reference model.py:264-271)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    x = torch.randn(batch, 4, input_size, input_size, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    y = 0.5 * torch.randn(batch, hidden, generator=g)
    y2 = torch.randn(batch, tokens, hidden, generator=g)
    w = torch.sigmoid(0.8 * torch.randn(batch, tokens, 1, generator=g))
    to = lambda a: a.to(device=device, dtype=dtype)
    return {"x": to(x), "t": t.to(device), "y": to(y), "y2": to(y2), "w": to(w)}
