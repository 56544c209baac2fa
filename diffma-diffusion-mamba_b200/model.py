"""``DiffMa`` and the ``DiffMa_models`` registry with the reference's interface (row a11 of SURVEY.md section 8).

``DiffMa_models[name](input_size=, dt_rank=, d_state=, use_mamba2=)`` and
``DiffMa.forward(x, t, y, y2, w)`` follow reference model.py:112-316 / :377-673 -- same constructor kwargs, same
sub-module and parameter names (``x_embedder.proj``, ``t_embedder.mlp.{0,2}``, ``pos_embed``, ``blocks.N...``,
``final_layer.{linear,adaLN_modulation.1}``), so a reference checkpoint's ``state_dict`` loads with
``strict=True``.  The GPU box has no ``/root/reference``; this mirror is what bench.py and the -m gpu tests run.
The reference's own ``model.py`` runs unchanged on top of ``shims/`` as well.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
from torch import nn

from . import scan_orders
from .blocks import (DiTBlock, EfficientVMamba_MambaBlock, Spiral_MambaBlock, ViM_MambaBlock, VMamba_MambaBlock,
                     Zig_MambaBlock, modulate)


_STEP_HEAD = os.environ.get("DIFFMA_STEP_HEAD", "1") != "0"     # inference: fused patch embed + conditioning (dm_step_head)

class PatchEmbed(nn.Module):
    def __init__(self, img_size=28, patch_size=2, stride=2, in_chans=4, embed_dim=512, norm_layer=None, flatten=True):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = ((img_size - patch_size) // stride + 1, (img_size - patch_size) // stride + 1)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=stride)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class TimestepEmbed(nn.Module):
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size
        self._freqs = {}

    def timestep_embedding(self, t, dim, max_period=10000):
        half = dim // 2
        key = (str(t.device), half)
        freqs = self._freqs.get(key)
        if freqs is None:   # built once per device: no per-step host->device upload (reference model.py:74-76 does one)
            freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
            self._freqs[key] = freqs
        args = t[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        if dim % 2:
            emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
        return emb

    def forward(self, t):
        return self.mlp(self.timestep_embedding(t, self.frequency_embedding_size))


class FinalLayer(nn.Module):
    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size * 2, 2 * hidden_size, bias=True))

    def forward(self, x, c):
        shift, scale = self.adaLN_modulation(c).chunk(2, dim=1)
        return self.linear(modulate(self.norm_final(x), shift, scale))


def get_2d_sincos_pos_embed(embed_dim, grid_size):
    """MAE sin-cos table as wired at reference model.py:325-372 (meshgrid with w first)."""
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 4, dtype=np.float64) / (embed_dim / 4.0))
    gh, gw = np.meshgrid(np.arange(grid_size, dtype=np.float32), np.arange(grid_size, dtype=np.float32), indexing="ij")

    def emb(p):
        o = np.einsum("m,d->md", p.reshape(-1), omega)
        return np.concatenate([np.sin(o), np.cos(o)], axis=1)

    return np.concatenate([emb(gw), emb(gh)], axis=1)


class DiffMa(nn.Module):
    def __init__(self, input_size=28, patch_size=2, strip_size=2, in_channels=4, hidden_size=512, depth=16,
                 learn_sigma=True, block_type="spiral", dt_rank=16, d_state=16, use_mamba2=False):
        super().__init__()
        self.learn_sigma, self.depth, self.in_channels = learn_sigma, depth, in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.patch_size, self.input_size, self.block_type = patch_size, input_size, block_type
        self.x_embedder = PatchEmbed(input_size, patch_size, strip_size, in_channels, hidden_size)
        self.t_embedder = TimestepEmbed(hidden_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.x_embedder.num_patches, hidden_size), requires_grad=False)
        n = int(input_size / patch_size)
        common = dict(D_dim=hidden_size, E_dim=hidden_size * 2, dim_inner=hidden_size * 2, dt_rank=dt_rank,
                      d_state=d_state, use_mamba2=use_mamba2)
        if block_type == "spiral":
            ml, inv = scan_orders.spiral(n)
            self.blocks = nn.ModuleList([
                Spiral_MambaBlock(token_list=ml[(2 * i) % len(ml)], token_list_reversal=ml[(2 * i) % len(ml) + 1],
                                  origina_list=inv[(2 * i) % len(ml)], origina_list_reversal=inv[(2 * i) % len(ml) + 1],
                                  **common) for i in range(depth)])
        elif block_type == "zig":
            self.blocks = nn.ModuleList([
                Zig_MambaBlock(token_list=scan_orders.zig(n, i)[0], origina_list=scan_orders.zig(n, i)[1], **common)
                for i in range(depth)])
        elif block_type == "vim":
            self.blocks = nn.ModuleList([ViM_MambaBlock(**common) for _ in range(depth)])
        elif block_type == "vmamba":
            ol, il = scan_orders.vmamba_(n)
            self.blocks = nn.ModuleList([VMamba_MambaBlock(token_list=ol, origina_list=il, **common)
                                         for _ in range(depth)])
        elif block_type == "efficientVMamba":
            self.blocks = nn.ModuleList([EfficientVMamba_MambaBlock(**common) for _ in range(depth)])
        elif block_type == "DiT":
            self.blocks = nn.ModuleList([DiTBlock(hidden_size=hidden_size, num_heads=8) for _ in range(depth)])
        else:
            raise ValueError(block_type)
        self.final_layer = FinalLayer(hidden_size, patch_size, self.out_channels)
        self.initialize_weights()

    def initialize_weights(self):
        def _basic_init(module):
            if isinstance(module, nn.Linear):
                torch.nn.init.xavier_uniform_(module.weight)
                if module.bias is not None:
                    nn.init.constant_(module.bias, 0)
        self.apply(_basic_init)
        pe = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], int(self.x_embedder.num_patches ** 0.5))
        self.pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.x_embedder.proj.bias, 0)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for block in self.blocks:
            nn.init.constant_(block.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(block.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].weight, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    def unpatchify(self, x):
        c, p = self.out_channels, self.x_embedder.patch_size[0]
        h = w = int(x.shape[1] ** 0.5)
        assert h * w == x.shape[1]
        x = x.reshape(x.shape[0], h, w, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, h * p, h * p)

    # ---- inference-only caches (keyed on parameter versions, rebuilt when weights change) ----------------
    def _cached(self, name, params, build):
        from . import ops
        key = ops.weights_key(params, str(params[0].device))
        slot = self.__dict__.setdefault("_icache", {})
        if name not in slot or slot[name][0] != key:
            with torch.no_grad():
                slot[name] = (key, build())
        return slot[name][1]

    def _fused_mods(self, c, act, silu_c=None):
        """adaLN of EVERY block and of the final layer in one GEMM: c is the same for all of them in a forward (the
        reference recomputes it per block, block/mamba_block.py:101).  Returns blocks (B, depth, 3D) and final (B, 2D),
        fp32; weights cached in the act dtype.  ``silu_c``: silu(c) already in the act dtype (``ops.step_head``)."""
        import torch.nn.functional as F
        lins = [blk.adaLN_modulation[1] for blk in self.blocks] + [self.final_layer.adaLN_modulation[1]]
        ps = [p for lin in lins for p in (lin.weight, lin.bias)]
        w, b = self._cached(f"ada_{act}", ps, lambda: (torch.cat([lin.weight for lin in lins]).to(act).contiguous(),
                                                        torch.cat([lin.bias for lin in lins]).to(act).contiguous()))
        with torch.autocast("cuda", enabled=False):
            mods = F.linear(F.silu(c.float()).to(act) if silu_c is None else silu_c, w, b).float()
        D = self.pos_embed.shape[-1]
        nb = self.depth * 3 * D
        return mods[:, :nb].view(mods.shape[0], self.depth, 3 * D), mods[:, nb:]

    def _t_embedding(self, t):
        """t_embedder(t) for integer timesteps from a (1000, D) table built once per weights version: the embedder is a
        pure function of the integer step (reference model.py:49-85 recomputes cos/sin + 2 Linears every call)."""
        if t.dtype not in (torch.int64, torch.int32):
            return self.t_embedder(t)
        ps = list(self.t_embedder.parameters())
        rows = int(getattr(self, "t_table_rows", 1000))     # number of ORIGINAL diffusion steps (create_model_and_diffusion
        #                                                     sets it; the reference's scripts always use 1000)

        def build():
            with torch.autocast("cuda", enabled=False):     # fp32 table whatever mode the first caller runs in
                return self.t_embedder(torch.arange(rows, device=ps[0].device)).float()
        table = self._cached(f"temb{rows}", ps, build)
        # an index beyond the table would be a device-side assert far from here: clamp is wrong, so fail loudly instead
        # when the host can see it (cheap: only for CPU-resident or tiny checks is this synchronising -- it is not done
        # on CUDA tensors; callers with more steps set ``t_table_rows``)
        if not t.is_cuda and int(t.max()) >= rows:
            raise IndexError(f"timestep {int(t.max())} beyond the {rows}-row embedding table; set model.t_table_rows")
        return table.index_select(0, t.long())

    def _patch_tables(self):
        """(unfolded conv weight (C*p*p, D) fp32, pos_embed + conv bias (L, D) fp32), cached per weights version."""
        return self._cached("patch", [self.x_embedder.proj.weight, self.x_embedder.proj.bias, self.pos_embed],
                            lambda: (self.x_embedder.proj.weight.reshape(self.x_embedder.proj.weight.shape[0], -1).t().contiguous().float(),
                                     (self.pos_embed[0] + self.x_embedder.proj.bias[None, :]).float().contiguous()))

    def _t_table(self):
        """TimestepEmbedder at the integer steps 0 .. t_table_rows - 1 (see ``_t_embedding``), fp32 (rows, D)."""
        ps = list(self.t_embedder.parameters())
        rows = int(getattr(self, "t_table_rows", 1000))

        def build():
            with torch.autocast("cuda", enabled=False):
                return self.t_embedder(torch.arange(rows, device=ps[0].device)).float().contiguous()
        return self._cached(f"temb{rows}", ps, build)

    def _embed_patches(self, x):
        """PatchEmbed conv (kernel = stride = patch) as one matmul on unfolded patches + (bias + pos_embed) table."""
        p = self.patch_size
        N, C, H, W = x.shape
        g = H // p
        wb = self._patch_tables()
        patches = x.float().view(N, C, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(N, g * g, C * p * p)
        with torch.autocast("cuda", enabled=False):
            return torch.baddbmm(wb[1].unsqueeze(0), patches, wb[0].unsqueeze(0).expand(N, -1, -1))

    def forward(self, x, t, y, y2, w):
        """x (N,C,H,W), t (N,), y (N,D), y2 (N,T,D) [or already pooled (N,D)], w (N,T,1) -> (N, out_channels, H, W)
        [reference model.py:264-301]"""
        fused = ((not torch.is_grad_enabled()) and x.is_cuda and self.block_type == "spiral"
                 and self.pos_embed.shape[-1] == 512 and self.x_embedder.proj.stride[0] == self.patch_size)
        if fused:
            from . import ops
            from .mixer import _act_dtype
            act = _act_dtype(x)
            if (_STEP_HEAD and t.dtype == torch.int64 and x.dtype == torch.float32 and x.is_contiguous()
                    and y.dtype == torch.float32 and y2.dtype == torch.float32 and x.shape[2] == x.shape[3]
                    and x.shape[1] * self.patch_size ** 2 * 16 * 4 <= 48 * 1024):
                y2m = y2                               # pooled (N, D) or not (N, T, D): the kernel takes the token mean itself
                # patch embedding + positional table + conditioning vector (timestep table row + y, + pooled y2) + SiLU in one
                # launch (dm_step_head) instead of eight
                wb = self._patch_tables()
                h, sc = ops.step_head(x, wb[0], wb[1], self.patch_size, t, self._t_table(), y.contiguous(), y2m.contiguous(), act)
                mods, fmod = self._fused_mods(None, act, silu_c=sc)
            else:
                y2m = y2 if y2.dim() == 2 else torch.mean(y2, dim=1)
                h = self._embed_patches(x)
                te = self._t_embedding(t)
                c = torch.cat((te + y, te + y2m), dim=1)
                mods, fmod = self._fused_mods(c, act)
            B, L, D = h.shape
            ones, zeros = self._cached("fl_affine", [self.pos_embed], lambda: (torch.ones(D, device=h.device),
                                                                               torch.zeros(D, device=h.device)))
            wrow = None if w is None else w.reshape(B * L).float().contiguous()
            # One row kernel opens block 0; after that ``spiral_post_mix_pre`` closes block i (sigmoid mix + gated
            # residual, reference mamba_block.py:111-114) and opens block i+1 (long-skip add of model.py:290-292,
            # LayerNorm, adaLN modulate, soft mask: mamba_block.py:101-105) -- or the final layer's LN + modulate
            # (eps 1e-6, no affine, model.py:92-109) -- in one pass over the residual stream.
            outs = []
            skip = None
            W0 = self.blocks[0]._fused_weights(act)
            x2 = ops.spiral_pre(h, None, W0["ln1"][0], W0["ln1"][1], mods[:, 0], wrow, act)
            for i in range(self.depth):
                blk = self.blocks[i]
                Wb = blk._fused_weights(act)
                ab, hidden = blk._fused_core(x2, B, L, act)
                if i + 1 < self.depth:
                    j = i + 1
                    skip_next = outs[self.depth - j - 1] if (j > self.depth / 2) else None
                    Wn = self.blocks[j]._fused_weights(act)
                    h, x2 = blk._post_mix(h, skip, ab, hidden, Wb, mods[:, i],
                                          pre=(skip_next, Wn["ln1"][0], Wn["ln1"][1], mods[:, j], wrow, 1e-5))
                    skip = skip_next
                else:
                    h, x2 = blk._post_mix(h, skip, ab, hidden, Wb, mods[:, i], pre=(None, ones, zeros, fmod, None, 1e-6))
                outs.append(h)
            hn = x2[0].view(h.shape)
            with torch.autocast("cuda", enabled=False):
                lw, lb = self._cached(f"fl_lin_{act}", [self.final_layer.linear.weight, self.final_layer.linear.bias],
                                      lambda: (self.final_layer.linear.weight.to(act).contiguous(),
                                               self.final_layer.linear.bias.to(act).contiguous()))
                if _STEP_HEAD:      # FinalLayer.linear + unpatchify in one launch (bf16, patch^2 * out_channels <= 128)
                    img = ops.final_linear_unpatchify(x2[0], lw, lb, B, self.patch_size, self.out_channels)
                    if img is not None:
                        return img
                o = torch.nn.functional.linear(hn, lw, lb)
            return self.unpatchify(o)
        x = self.x_embedder(x) + self.pos_embed
        t = self.t_embedder(t)
        y2m = y2 if y2.dim() == 2 else torch.mean(y2, dim=1)
        c = torch.cat((t + y, t + y2m), dim=1)
        outs = []
        for i in range(self.depth):
            if i == 0:
                x = self.blocks[i](x, c, w)
            elif i > self.depth / 2:
                x = self.blocks[i](outs[-1] + outs[self.depth - i - 1], c, w)
            else:
                x = self.blocks[i](outs[-1], c, w)
            outs.append(x)
        return self.unpatchify(self.final_layer(x, c))

    def forward_with_cfg(self, x, t, y, y2, w, cfg_scale):
        half = x[: len(x) // 2]
        combined = torch.cat([half, half], dim=0)
        out = self.forward(combined, t, y, y2, w)
        eps, rest = out[:, :3], out[:, 3:]
        cond, uncond = torch.split(eps, len(eps) // 2, dim=0)
        half_eps = uncond + cfg_scale * (cond - uncond)
        return torch.cat([torch.cat([half_eps, half_eps], dim=0), rest], dim=1)


_DEPTH = {"XXL": 56, "XL": 28, "L": 16, "B": 8, "S": 4, "BL": 13, "SB": 7}
_FAMILY = {"DiffMa": "spiral", "ZigMa": "zig", "ViM": "vim", "VMamba": "vmamba", "EMamba": "efficientVMamba",
           "DiT": "DiT"}


def _factory(depth, patch, block_type):
    def make(**kwargs):
        return DiffMa(depth=depth, hidden_size=512, patch_size=patch, strip_size=patch, block_type=block_type, **kwargs)
    return make


def _registry():
    reg = {}
    sizes = {"DiffMa": ("XXL", "XL", "L", "B", "S"), "ZigMa": ("XL", "L", "B", "S"), "ViM": ("XL", "L", "B", "S"),
             "VMamba": ("XL", "L", "B", "S"), "EMamba": ("XL", "L", "B", "S"), "DiT": ("XL", "L", "B", "S")}
    for fam, szs in sizes.items():
        for s in szs:
            for p in (2, 4, 7):
                reg[f"{fam}-{s}/{p}"] = _factory(_DEPTH[s], p, _FAMILY[fam])
    for fam in ("ZigMa", "ViM", "VMamba", "EMamba"):
        reg[f"{fam}-BL/2"] = _factory(_DEPTH["BL"], 2, _FAMILY[fam])
    reg["DiT-SB/2"] = _factory(_DEPTH["SB"], 2, "DiT")
    return reg


DiffMa_models = _registry()
