"""Functional CPU restatement of the callers around the external ops (TEST INFRASTRUCTURE).

Orchestration parity is PINNED: ``tests/golden/model_*.npz`` hold outputs of the reference's own
``model.py`` / ``block/*.py`` (run unmodified on CPU over ``oracle/ref_shims``), and
``tests/test_oracle_golden.py`` checks this file against them.  The arithmetic of the inner ops
comes from ``ref_ops`` (parity unpinned, see there).

Everything works on a plain ``state_dict`` whose keys are the reference's parameter names, so the
same function checks the reference model, and the product's mirror modules.

Follows: ``Mamba.forward`` block/mamba.py:317-403, ``Mamba2.forward`` block/mamba2.py:359-712,
block ``forward``s block/mamba_block.py:100-115,185-192,246-253,319-326,381-388,414-418,
``DiffMa.forward`` model.py:264-301, ``CT_Encoder.forward`` block/CT_encoder.py:37-44.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_ops, ref_scan_orders


# ------------------------------------------------------------------------------------------------
# direction handling (reference-owned semantics)
# ------------------------------------------------------------------------------------------------
def _eff_orders(L):
    """Token index lists of the four EfficientVMamba sub-scans (block/mamba.py:170-183)."""
    s = int(round(math.sqrt(L)))
    assert s * s == L and s % 2 == 0, "EfficientVMamba cross-scan needs an even square grid"
    grid = np.arange(L).reshape(s, s)
    return [grid[::2, ::2].reshape(-1).tolist(), grid.T[::2, 1::2].reshape(-1).tolist(),
            grid[::2, 1::2].reshape(-1).tolist(), grid.T[1::2, 1::2].reshape(-1).tolist()]


def _directions(scan_type, L, orders):
    """-> list of (gather index list or None)."""
    if scan_type == "spiral":
        return [None, orders["token_list"], orders["token_list_reversal"]]
    if scan_type == "zigma":
        return [orders["token_list"]]
    if scan_type == "vmamba":
        return list(orders["token_list"])
    if scan_type == "eff":
        return _eff_orders(L)
    raise ValueError(scan_type)


def _mix(scan_type, h, inner, orders, vim_flip_dim):
    """Apply ``inner`` ((B,Lk,d_in_proj tokens-major) -> (B,Lk,d_model)) per direction and merge."""
    B, L, _ = h.shape
    if scan_type == "vim":
        out1 = inner(h)
        out2 = inner(torch.flip(h, [1]))
        # Mamba-1 flips the FEATURE axis of out2 (block/mamba.py:366);
        # Mamba-2 flips the token axis (block/mamba2.py:522).
        return (out1 + torch.flip(out2, [vim_flip_dim])) / 2
    out = None
    for idx in _directions(scan_type, L, orders):
        if idx is None:
            o = inner(h)
            out = o if out is None else out + o
            continue
        idx_t = torch.as_tensor(idx, dtype=torch.long)
        o = inner(h[:, idx_t, :])
        if out is None:
            out = torch.zeros(B, L, o.shape[-1], dtype=o.dtype)
        # merge: y[:, i] += o[:, inv[i]]  <=>  y[:, idx[j]] += o[:, j]
        out.index_add_(1, idx_t, o)
    return out


def mamba1_mixer_ref(sd, prefix, h, scan_type, orders=None, compute_dtype=torch.float32):
    """``Mamba.forward(hidden_states, scan_type)`` (block/mamba.py:317-403)."""
    p = lambda n: sd[prefix + n].to(compute_dtype)
    dtype_in = h.dtype
    xz = h.to(compute_dtype) @ p("in_proj.weight").t()                      # (B,L,2D) tokens-major
    A = -torch.exp(p("A_log"))

    def inner(xz_k):
        return ref_ops.mamba_inner_ref(xz_k.transpose(1, 2), p("conv1d.weight"), p("conv1d.bias"),
                                       p("x_proj.weight"), p("dt_proj.weight"), p("out_proj.weight"), None,
                                       A, None, None, p("D"), delta_bias=p("dt_proj.bias"),
                                       delta_softplus=True, compute_dtype=compute_dtype)

    return _mix(scan_type, xz, inner, orders, vim_flip_dim=2).to(dtype_in)


def mamba2_mixer_ref(sd, prefix, u, scan_type, orders=None, headdim=64, chunk_size=256,
                     compute_dtype=torch.float32):
    """``Mamba2.forward(u, scan_type)`` (block/mamba2.py:359-712), ngroups=1, rmsnorm, norm_before_gate=False."""
    p = lambda n: sd[prefix + n].to(compute_dtype)
    dtype_in = u.dtype
    zxbcdt = u.to(compute_dtype) @ p("in_proj.weight").t()
    A = -torch.exp(p("A_log"))

    def inner(z_k):
        return ref_ops.mamba_split_conv1d_scan_ref(
            z_k, p("conv1d.weight").reshape(p("conv1d.weight").shape[0], -1), p("conv1d.bias"), p("dt_bias"), A,
            p("D"), chunk_size, activation="silu", rmsnorm_weight=p("norm.weight"), rmsnorm_eps=1e-5,
            outproj_weight=p("out_proj.weight"), outproj_bias=None, headdim=headdim, ngroups=1,
            norm_before_gate=False, compute_dtype=compute_dtype)

    return _mix(scan_type, zxbcdt, inner, orders, vim_flip_dim=1).to(dtype_in)


# ------------------------------------------------------------------------------------------------
# blocks
# ------------------------------------------------------------------------------------------------
def _modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def _linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype))


def _layer_norm(sd, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype), eps)


_SCAN_OF_BLOCK = {"spiral": "spiral", "zig": "zigma", "vim": "vim", "vmamba": "vmamba", "efficientVMamba": "eff"}


def block_ref(sd, prefix, block_type, x, c, w, orders, use_mamba2, compute_dtype=torch.float32):
    """All Mamba block flavours of block/mamba_block.py; ``orders`` as produced by ``block_orders``."""
    x = x.to(compute_dtype)
    mixer = mamba2_mixer_ref if use_mamba2 else mamba1_mixer_ref
    scan = _SCAN_OF_BLOCK[block_type]
    shift, scale, gate = _linear(sd, prefix + "adaLN_modulation.1", F.silu(c.to(compute_dtype))).chunk(3, dim=1)
    x_ssm = _modulate(_layer_norm(sd, prefix + "norm1", x), shift, scale)
    if block_type == "spiral":
        a = mixer(sd, prefix + "mamba1.", x_ssm, scan, orders, compute_dtype=compute_dtype)
        b = mixer(sd, prefix + "mamba2.", x_ssm * w.to(compute_dtype), scan, orders, compute_dtype=compute_dtype)
        hcat = _layer_norm(sd, prefix + "attention_network.0", torch.cat([a, b], dim=-1))
        alpha = torch.sigmoid(_linear(sd, prefix + "attention_network.3",
                                      F.silu(_linear(sd, prefix + "attention_network.1", hcat))))
        mixed = alpha * a + (1 - alpha) * b
    else:
        mixed = mixer(sd, prefix + "mamba.", x_ssm, scan, orders, compute_dtype=compute_dtype)
    return x + gate.unsqueeze(1) * mixed


def block_orders(block_type, grid, i):
    """Scan-order kwargs of block ``i`` exactly as model.py:144-194 wires them."""
    if block_type == "spiral":
        ml, inv = ref_scan_orders.spiral(grid)
        k = (2 * i) % len(ml)
        return {"token_list": ml[k], "token_list_reversal": ml[k + 1],
                "origina_list": inv[k], "origina_list_reversal": inv[k + 1]}
    if block_type == "zig":
        o, inv = ref_scan_orders.zig(grid, i)
        return {"token_list": o, "origina_list": inv}
    if block_type == "vmamba":
        o, inv = ref_scan_orders.vmamba_(grid)
        return {"token_list": o, "origina_list": inv}
    return {}


# ------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim=256, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def diffma_forward_ref(sd, cfg, x, t, y, y2, w, compute_dtype=torch.float32):
    """``DiffMa.forward`` (model.py:264-301). cfg: depth, patch_size, block_type, use_mamba2."""
    depth, patch, block_type = cfg["depth"], cfg["patch_size"], cfg["block_type"]
    cd = compute_dtype
    xe = F.conv2d(x.to(cd), sd["x_embedder.proj.weight"].to(cd), sd["x_embedder.proj.bias"].to(cd), stride=patch)
    grid = xe.shape[-1]
    h = xe.flatten(2).transpose(1, 2) + sd["pos_embed"].to(cd)
    te = timestep_embedding(t).to(cd)
    te = _linear(sd, "t_embedder.mlp.2", F.silu(_linear(sd, "t_embedder.mlp.0", te)))
    c = torch.cat((te + y.to(cd), te + y2.to(cd).mean(dim=1)), dim=1)
    outs = []
    for i in range(depth):
        if i == 0:
            inp = h
        elif i > depth / 2:
            inp = outs[-1] + outs[depth - i - 1]            # long skip, model.py:290-292
        else:
            inp = outs[-1]
        outs.append(block_ref(sd, f"blocks.{i}.", block_type, inp, c, w, block_orders(block_type, grid, i),
                              cfg.get("use_mamba2", False), compute_dtype=cd))
    h = outs[-1]
    shift, scale = _linear(sd, "final_layer.adaLN_modulation.1", F.silu(c)).chunk(2, dim=1)
    h = _modulate(F.layer_norm(h, (h.shape[-1],), eps=1e-6), shift, scale)
    h = _linear(sd, "final_layer.linear", h)
    # unpatchify, model.py:249-262
    n, T, _ = h.shape
    co = h.shape[-1] // (patch * patch)
    g = int(round(math.sqrt(T)))
    h = h.reshape(n, g, g, patch, patch, co)
    return torch.einsum("nhwpqc->nchpwq", h).reshape(n, co, g * patch, g * patch)


def ct_encoder_ref(sd, x, patch_size=2, compute_dtype=torch.float32):
    """``CT_Encoder.forward`` (block/CT_encoder.py:37-44) -> (weight (N,T,1), y2 (N,T,D))."""
    cd = compute_dtype
    e = F.conv2d(x.to(cd), sd["vision_embedding.proj.weight"].to(cd), sd["vision_embedding.proj.bias"].to(cd),
                 stride=patch_size).flatten(2).transpose(1, 2)

    def fc(v):
        return _linear(sd, "fc.2", F.relu(_linear(sd, "fc.0", v)))

    weight = torch.sigmoid(fc(e.mean(dim=-1)) + fc(e.amax(dim=-1))).unsqueeze(-1)
    return weight, _layer_norm(sd, "norm", e * weight)
